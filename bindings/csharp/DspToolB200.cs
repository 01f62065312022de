// bindings/csharp/DspToolB200.cs — DspToolB200 : IDspTool (src/VGAudio.Tools/GcAdpcm/IDspTool.cs:5-12).
//
// With this class the reference's own differential harness (`VGAudio.Tools gcadpcm`, src/VGAudio.Tools/GcAdpcm/Encode.cs:44-150,
// which compares DspToolVGAudio against Nintendo's DLL through DspToolDll) runs unchanged against libvgaudio_b200.so:
// add the file to src/VGAudio.Tools/GcAdpcm/ and hand `new DspToolB200()` to the comparison in place of DspToolDll.
// It mirrors DspToolVGAudio.cs:6-36 call for call; the cdecl precedent is DspToolDll.cs:15-29 (correlateCoefs / encodeFrame /
// encode / decode over ADPCMINFO, Native.cs:18-32) - the ADPCMINFO encode / decode pair maps onto the one-channel forms of
// vgb_gcadpcm_encode_batch / vgb_gcadpcm_decode_batch below.  NOT compiled in this repository (no .NET toolchain in the image).
using System;
using VGAudio.Codecs.GcAdpcm;
using VGAudio.Formats.GcAdpcm;
using VGAudio.Native;

namespace VGAudio.Tools.GcAdpcm
{
    public unsafe class DspToolB200 : IDspTool
    {
        public DspToolB200(int device = 0) => VgAudioB200.Check(VgAudioB200.vgb_init(device, 0));

        // DspToolVGAudio.EncodeChannel: CalculateCoefficients + GcAdpcmEncoder.Encode (GcAdpcmFormat.cs:129-135)
        public GcAdpcmChannel EncodeChannel(short[] pcm)
        {
            int sampleCount = pcm.Length;
            var coefs = new short[16];
            var adpcm = new byte[GcAdpcmMath.SampleCountToByteCount(sampleCount)];
            fixed (short* p = pcm, c = coefs)
            fixed (byte* a = adpcm)
            {
                short* pp = p;
                byte* aa = a;
                VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_encode_batch(&pp, &sampleCount, null, null, 1, c, &aa, null, IntPtr.Zero));
            }
            return new GcAdpcmChannel(adpcm, coefs, sampleCount);
        }

        // GcAdpcmCoefficients.CalculateCoefficients (GcAdpcmCoefficients.cs:9-110); the DLL's correlateCoefs
        public short[] DspCorrelateCoefs(short[] pcm)
        {
            var coefs = new short[16];
            int n = pcm.Length;
            fixed (short* p = pcm, c = coefs)
            {
                short* pp = p;
                VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_coefs_batch(&pp, &n, 1, c));
            }
            return coefs;
        }

        // GcAdpcmEncoder.DspEncodeFrame (GcAdpcmEncoder.cs:48-94): pcmInOut = two history samples + 14, rewritten with the
        // reconstruction; the DLL's encodeFrame
        public void DspEncodeFrame(short[] pcmInOut, int sampleCount, byte[] adpcmOut, short[] coefsIn)
        {
            fixed (short* p = pcmInOut, c = coefsIn)
            fixed (byte* a = adpcmOut)
                VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_encode_frames(p, &sampleCount, c, 1, a));
        }

        public short[] DecodeChannel(GcAdpcmChannel channel) =>
            DecodeAdpcm(channel.GetAdpcmAudio(), channel.Coefs, channel.SampleCount);

        // GcAdpcmDecoder.Decode(adpcm, coefs, new GcAdpcmParameters { SampleCount = sampleCount }) (GcAdpcmDecoder.cs:10-54)
        public short[] DecodeAdpcm(byte[] adpcm, short[] coefs, int sampleCount)
        {
            var pcm = new short[sampleCount];
            int nBytes = adpcm.Length;
            var cfg = new VgbGcParams { SampleCount = sampleCount, History1 = 0, History2 = 0 };
            fixed (byte* a = adpcm)
            fixed (short* c = coefs, o = pcm)
            {
                byte* aa = a;
                short* oo = o;
                VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_decode_batch(&aa, &nBytes, c, &cfg, 1, &oo));
            }
            return pcm;
        }
    }
}
