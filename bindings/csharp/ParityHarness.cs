// bindings/csharp/ParityHarness.cs — closes the "parity unpinned" gap of DESIGN.md §3 on a machine that has .NET:
// encodes the seeded synthetic channels with the UNMODIFIED managed path and with libvgaudio_b200.so and compares
// coefficients and ADPCM bytes.  Modelled on the reference's own differential tool
// (src/VGAudio.Tools/GcAdpcm/Encode.cs:44-150, which compares VGAudio against Nintendo's dsptool DLL).
// Build: add to a console project referencing src/VGAudio/VGAudio.csproj.  NOT compiled here.
using System;
using System.Linq;
using VGAudio.Codecs.GcAdpcm;
using VGAudio.Native;

internal static unsafe class ParityHarness
{
    private static short[] Sine(int n, double f, int rate) =>
        Enumerable.Range(0, n).Select(i => (short)(short.MaxValue * Math.Sin(2 * Math.PI * f / rate * i))).ToArray();

    private static int Main()
    {
        double[] freqs = { 261.63, 329.63, 392, 523.25, 659.25, 783.99, 1046.50, 130.81 };   // GenerateAudio.cs:14
        int bad = 0;
        foreach (double f in freqs)
        {
            short[] pcm = Sine(48000, f, 48000);
            short[] coefsManaged = GcAdpcmCoefficients.CalculateCoefficients(pcm);
            byte[] adpcmManaged = GcAdpcmEncoder.Encode(pcm, coefsManaged);

            var coefsNative = new short[16];
            var adpcmNative = new byte[adpcmManaged.Length];
            int len = pcm.Length;
            fixed (short* p = pcm) fixed (short* c = coefsNative) fixed (byte* a = adpcmNative)
            {
                short* pp = p; byte* aa = a;
                VgAudioB200.Check(VgAudioB200.vgb_gcadpcm_encode_batch(&pp, &len, null, null, 1, c, &aa, null, IntPtr.Zero));
            }
            bool same = coefsManaged.SequenceEqual(coefsNative) && adpcmManaged.SequenceEqual(adpcmNative);
            Console.WriteLine($"{f,8:F2} Hz: {(same ? "identical" : "DIFFERENT")}");
            if (!same) bad++;
        }
        return bad;
    }
}
