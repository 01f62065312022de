// bindings/csharp/GcAdpcmFormat.B200.cs — drop-in bodies for the two hot methods of
// src/VGAudio/Formats/GcAdpcm/GcAdpcmFormat.cs.  Everything around them (builders, loop handling, containers) is
// untouched: the Parallel.For over channels (GcAdpcmFormat.cs:65-68 and :45-48) becomes ONE batched native call.
// NOT compiled here (no .NET toolchain in the build image).
using System;
using System.Runtime.InteropServices;
using VGAudio.Codecs.GcAdpcm;
using VGAudio.Formats.Pcm16;
using VGAudio.Native;
using static VGAudio.Codecs.GcAdpcm.GcAdpcmMath;

namespace VGAudio.Formats.GcAdpcm
{
    public partial class GcAdpcmFormat
    {
        // replaces GcAdpcmFormat.EncodeFromPcm16(Pcm16Format, GcAdpcmParameters)  (GcAdpcmFormat.cs:58-74)
        public override unsafe GcAdpcmFormat EncodeFromPcm16(Pcm16Format pcm16, GcAdpcmParameters config)
        {
            int n = pcm16.ChannelCount;
            var channels = new GcAdpcmChannel[n];
            int frameCount = pcm16.SampleCount.DivideByRoundUp(14) * n;
            config?.Progress?.SetTotal(frameCount);                                  // :62-63

            int sampleCount = config == null || config.SampleCount == -1 ? pcm16.SampleCount : config.SampleCount;
            var coefs = new short[n * 16];
            var adpcm = new byte[n][];
            var pins = new GCHandle[2 * n];                                           // short[][] / byte[][] are not blittable
            var pcmPtr = stackalloc short*[n];
            var outPtr = stackalloc byte*[n];
            var lens = stackalloc int[n];
            var prm = stackalloc VgbGcParams[n];
            VgbProgress cb = config?.Progress == null ? null : (u, d) => config.Progress.ReportAdd((int)d);
            try
            {
                for (int i = 0; i < n; i++)
                {
                    adpcm[i] = new byte[SampleCountToByteCount(sampleCount)];         // GcAdpcmEncoder.cs:18
                    pins[2 * i] = GCHandle.Alloc(pcm16.Channels[i], GCHandleType.Pinned);
                    pins[2 * i + 1] = GCHandle.Alloc(adpcm[i], GCHandleType.Pinned);
                    pcmPtr[i] = (short*)pins[2 * i].AddrOfPinnedObject();
                    outPtr[i] = (byte*)pins[2 * i + 1].AddrOfPinnedObject();
                    lens[i] = pcm16.Channels[i].Length;
                    prm[i] = new VgbGcParams { SampleCount = config?.SampleCount ?? -1, History1 = config?.History1 ?? 0, History2 = config?.History2 ?? 0 };
                }
                fixed (short* c = coefs)
                    VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_encode_batch(pcmPtr, lens, prm, null, n, c, outPtr, cb, IntPtr.Zero));
            }
            finally { foreach (var h in pins) if (h.IsAllocated) h.Free(); }
            GC.KeepAlive(cb);

            for (int i = 0; i < n; i++)
            {
                var c = new short[16];
                Array.Copy(coefs, i * 16, c, 0, 16);
                channels[i] = new GcAdpcmChannel(adpcm[i], c, pcm16.SampleCount);     // EncodeChannel :134
            }
            return new GcAdpcmFormatBuilder(channels, pcm16.SampleRate)
                .WithLoop(pcm16.Looping, pcm16.LoopStart, pcm16.LoopEnd)
                .WithTracks(pcm16.Tracks)
                .Build();                                                             // :70-73 unchanged
        }

        // replaces GcAdpcmFormat.ToPcm16()  (GcAdpcmFormat.cs:42-54) for channels that need decoding
        public override unsafe Pcm16Format ToPcm16()
        {
            int n = Channels.Length;
            var pcm = new short[n][];
            var coefs = new short[n * 16];
            var pins = new GCHandle[2 * n];
            var inPtr = stackalloc byte*[n];
            var outPtr = stackalloc short*[n];
            var lens = stackalloc int[n];
            var prm = stackalloc VgbGcParams[n];
            try
            {
                for (int i = 0; i < n; i++)
                {
                    byte[] a = Channels[i].GetAdpcmAudio();
                    pcm[i] = new short[Channels[i].SampleCount];
                    Array.Copy(Channels[i].Coefs, 0, coefs, i * 16, 16);
                    pins[2 * i] = GCHandle.Alloc(a, GCHandleType.Pinned);
                    pins[2 * i + 1] = GCHandle.Alloc(pcm[i], GCHandleType.Pinned);
                    inPtr[i] = (byte*)pins[2 * i].AddrOfPinnedObject();
                    outPtr[i] = (short*)pins[2 * i + 1].AddrOfPinnedObject();
                    lens[i] = a.Length;
                    prm[i] = new VgbGcParams { SampleCount = Channels[i].SampleCount, History1 = Channels[i].StartContext.Hist1, History2 = Channels[i].StartContext.Hist2 };
                }
                fixed (short* c = coefs)
                    VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_decode_batch(inPtr, lens, c, prm, n, outPtr));
            }
            finally { foreach (var h in pins) if (h.IsAllocated) h.Free(); }
            return new Pcm16FormatBuilder(pcm, SampleRate).WithLoop(Looping, LoopStart, LoopEnd).WithTracks(Tracks).Build();
        }
    }
}
