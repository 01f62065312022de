// bindings/csharp/Containers.B200.cs — P/Invoke declarations and replacement bodies for the container layer either side of
// the codec path (include/vgaudio_b200.h, "Containers either side of the codec path"): WaveReader, DspWriter / DspReader,
// AdxWriter (+ CriAdxEncryption), HcaWriter (+ CriHcaEncryption) and the CLI's batch job.
// NOT compiled in this repository (no .NET toolchain in the build image); this is the file a VGAudio maintainer adds.
using System;
using System.Collections.Generic;
using System.IO;
using System.Linq;
using System.Runtime.InteropServices;

namespace VGAudio.Native
{
    [StructLayout(LayoutKind.Sequential)]
    internal struct VgbWaveInfo
    {
        public int ChannelCount, SampleRate, BitsPerSample, SampleCount, Looping, LoopStart, LoopEnd, Reserved;
        public long DataOffset, DataSize;
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct VgbDspDesc   // what DspWriter reads from GcAdpcmFormat + DspConfiguration (DspWriter.cs:17-36)
    {
        public int ChannelCount, SampleRate, SampleCount, Looping, LoopStart, LoopEnd;
        public int SamplesPerInterleave, LoopPointAlignment, NoTrim;   // 0 = 0x3800, 1, TrimFile = true
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct VgbAdxDesc   // what AdxWriter reads from CriAdxFormat + AdxConfiguration (AdxWriter.cs:18-55)
    {
        public int ChannelCount, SampleRate, SampleCount, Looping, LoopStart, LoopEnd, AlignmentSamples;
        public int FrameSize, Version, Type, HighpassFrequency, EncryptionType, NoTrim;
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct VgbAdxKey { public int Seed, Mult, Inc; }

    [StructLayout(LayoutKind.Sequential)]
    internal struct VgbConvertOptions
    {
        public int OutType;                 // 1 .dsp, 2 .adx, 3 .hca
        public int NoTrim;
        public int DspSamplesPerInterleave, DspLoopPointAlignment;
        public int AdxVersion, AdxFrameSize, AdxType, AdxFilterPlus1;
        public int AdxEncryptionType, AdxHasKey, AdxKeySeed, AdxKeyMult, AdxKeyInc;
        public int HcaQuality, HcaBitrate, HcaLimitBitrate;
        public int HcaKeyType;              // -1 = none (0 IS a key type)
        public int Reserved;
        public ulong HcaKeyCode;
        public long GroupBytes;
    }

    internal static unsafe class VgAudioB200Containers
    {
        private const string Lib = "vgaudio_b200";

        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int vgb_wave_parse(byte* file, long length, VgbWaveInfo* info);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int vgb_wave_read_batch(byte** files, long* lengths, VgbWaveInfo* info, int nFiles, short** pcmOut);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern long vgb_dsp_file_size(VgbDspDesc* desc);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_dsp_write_batch(VgbDspDesc* files, int nFiles, byte** adpcm, short* coefs, short* gain, short* startHist,
            short* loopContext, byte** filesOut);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern long vgb_adx_file_size(VgbAdxDesc* desc);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int vgb_adx_key_from_code(ulong keyCode, VgbAdxKey* key);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int vgb_adx_key_from_string([MarshalAs(UnmanagedType.LPStr)] string s, VgbAdxKey* key);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_adx_write_batch(VgbAdxDesc* files, int nFiles, byte** audio, int* audioLen, short* history, VgbAdxKey* key, byte** filesOut);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_adx_crypt_batch(byte** audio, int nChannels, int length, VgbAdxKey* key, int encryptionType, int frameSize);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int vgb_hca_key_tables(int keyType, ulong keyCode, byte* decrypt, byte* encrypt);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_hca_crypt_batch(byte** frames, int* frameCount, int nStreams, int frameSize, int keyType, ulong keyCode, int decrypt);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_hca_write_batch(VgbHcaInfo* info, int nFiles, byte** frames, int keyType, ulong keyCode, byte** comment, float* volume, byte** filesOut);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_convert_dsp_to_wave_batch(byte** files, long* lengths, int nFiles, long* outSizes, byte** filesOut, int* statusOut);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int vgb_convert_wave_batch(byte** files, long* lengths, int nFiles, VgbConvertOptions* options, long* outSizes,
            byte** filesOut, int* statusOut, VgbProgress progress, IntPtr user);
    }
}

namespace VGAudio.Containers.Dsp
{
    using VGAudio.Native;

    // Replacement body for DspWriter.WriteStream (Containers/Dsp/DspWriter.cs:42-52): the header fields and the
    // block interleave of WriteHeader / WriteData (:54-99) become one native call that returns the finished file.
    public partial class DspWriterB200
    {
        internal static unsafe byte[] GetFile(VGAudio.Formats.GcAdpcm.GcAdpcmFormat adpcm, DspConfiguration config)
        {
            var desc = new VgbDspDesc
            {
                ChannelCount = adpcm.ChannelCount, SampleRate = adpcm.SampleRate, SampleCount = adpcm.SampleCount,
                Looping = adpcm.Looping ? 1 : 0, LoopStart = adpcm.LoopStart, LoopEnd = adpcm.LoopEnd,
                SamplesPerInterleave = config.SamplesPerInterleave, LoopPointAlignment = config.LoopPointAlignment, NoTrim = config.TrimFile ? 0 : 1
            };
            long size = VgAudioB200Containers.vgb_dsp_file_size(&desc);
            if (size < 0) VgAudioB200.Check((int)size);
            var file = new byte[size];
            int n = adpcm.ChannelCount;
            byte[][] audio = adpcm.Channels.Select(c => c.GetAdpcmAudio()).ToArray();
            short[] coefs = adpcm.Channels.SelectMany(c => c.Coefs).ToArray();
            short[] gain = adpcm.Channels.Select(c => c.Gain).ToArray();
            short[] hist = adpcm.Channels.SelectMany(c => new[] { c.StartContext.Hist1, c.StartContext.Hist2 }).ToArray();
            short[] loop = adpcm.Channels.SelectMany(c => new[] { c.LoopContext.PredScale, c.LoopContext.Hist1, c.LoopContext.Hist2 }).ToArray();
            var pins = audio.Select(a => GCHandle.Alloc(a, GCHandleType.Pinned)).ToArray();
            try
            {
                byte** rows = stackalloc byte*[n];
                for (int i = 0; i < n; i++) rows[i] = (byte*)pins[i].AddrOfPinnedObject();
                fixed (byte* pf = file)
                fixed (short* pc = coefs, pg = gain, ph = hist, pl = loop)
                {
                    byte* outPtr = pf;
                    VgAudioB200.Check(VgAudioB200Containers.vgb_dsp_write_batch(&desc, 1, rows, pc, pg, ph, adpcm.Looping ? pl : null, &outPtr));
                }
            }
            finally { foreach (var h in pins) h.Free(); }
            return file;
        }
    }
}

namespace VGAudio.Cli
{
    using VGAudio.Native;

    // Replacement for the Parallel.ForEach of Batch.BatchConvert (src/VGAudio.Cli/Batch.cs:24-46) when every input is a
    // WAVE file and the output is .dsp / .adx / .hca: the managed side still enumerates, reads and writes files; a chunk of
    // file images goes through ONE native call (sizing pass, then the filling pass).
    internal static class BatchB200
    {
        public static unsafe void ConvertChunk(string[] inPaths, string[] outPaths, VgbConvertOptions options, Action<string> log, Action<int> reportAdd)
        {
            int n = inPaths.Length;
            byte[][] images = inPaths.Select(File.ReadAllBytes).ToArray();
            var inPins = images.Select(a => GCHandle.Alloc(a, GCHandleType.Pinned)).ToArray();
            var outPins = new List<GCHandle>();
            try
            {
                byte** inPtr = stackalloc byte*[n];
                byte** outPtr = stackalloc byte*[n];
                long* len = stackalloc long[n];
                long* outSize = stackalloc long[n];
                int* status = stackalloc int[n];
                for (int i = 0; i < n; i++) { inPtr[i] = (byte*)inPins[i].AddrOfPinnedObject(); len[i] = images[i].Length; }
                VgAudioB200.Check(VgAudioB200Containers.vgb_convert_wave_batch(inPtr, len, n, &options, outSize, null, status, null, IntPtr.Zero));
                var outputs = new byte[n][];
                for (int i = 0; i < n; i++)
                {
                    outPtr[i] = null;
                    if (status[i] != VgAudioB200.Ok) continue;
                    outputs[i] = new byte[outSize[i]];
                    outPins.Add(GCHandle.Alloc(outputs[i], GCHandleType.Pinned));
                    outPtr[i] = (byte*)outPins[outPins.Count - 1].AddrOfPinnedObject();
                }
                VgbProgress cb = (user, delta) => reportAdd((int)delta);   // progress.ReportAdd(1) per file (Batch.cs:45)
                VgAudioB200.Check(VgAudioB200Containers.vgb_convert_wave_batch(inPtr, len, n, &options, outSize, outPtr, status, cb, IntPtr.Zero));
                for (int i = 0; i < n; i++)
                {
                    if (status[i] != VgAudioB200.Ok) { log($"Error converting {Path.GetFileName(inPaths[i])}"); continue; }   // Batch.cs:39-43
                    Directory.CreateDirectory(Path.GetDirectoryName(outPaths[i]));
                    File.WriteAllBytes(outPaths[i], outputs[i]);
                }
            }
            finally
            {
                foreach (var h in inPins) h.Free();
                foreach (var h in outPins) h.Free();
            }
        }
    }
}
