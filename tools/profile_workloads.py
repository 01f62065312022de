#!/usr/bin/env python
"""tools/profile_workloads.py <workload> — runs ONE pass of a kernel family at a representative size, for `ncu` captures
(every library kernel of the workload launches once or twice; no warm-up loops, the profiler replays launches itself).

  c2      GC-ADPCM coefficients + time-parallel encode, 1024 ch x 30 s  (gc_coef_frames, gc_coef_refine, gc_encode<chain|run-on|cascade>)
  gcdec   GC-ADPCM decode of 2048 ch x 30 s + seek-table taps of 256 ch   (gc_decode<false>, gc_decode<true>)
  frames  DspEncodeFrame for 65536 independent frames                      (gc_encode_frames)
  adx     CRI ADX encode + decode, 1024 ch x 30 s                          (adx_encode<chain|run-on|cascade>, adx_decode)
  hca     CRI HCA encode + decode, 128 mono streams x 30 s, + MDCT taps     (hca_encode, hca_decode_parse/unpack/seam, hca_mdct128, hca_imdct128)
  ilv     block interleave / deinterleave, 512 x 2 x 822 864 B, vector and TMA variants
  ctn     the batch converter on 512 WAVE files: .dsp, keyed .adx, keyed .hca  (wave_split, dsp/adx/hca_assemble; capture with
          -k regex:"wave_split|assemble" so the codec kernels are not replayed)
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vgaudio_b200 as vg  # noqa: E402
from vgaudio_b200 import _native as N  # noqa: E402

RATE = 48000


def main():
    what = sys.argv[1]
    dev = torch.device("cuda", 0)
    N.check(vg.lib.vgb_init(0, 0))
    os.environ["VGB_PIPELINE_GROUPS"] = "1"  # one launch per kernel
    stream = torch.cuda.current_stream()
    n = 30 * RATE
    if what == "ctn":
        import bench_configs
        from vgaudio_b200 import containers as ct

        files, _ = bench_configs._batch_files(torch, bench, 512, dev)
        for opt in (ct.convert_options(ct.CONTAINER_DSP), ct.convert_options(ct.CONTAINER_ADX, adx_has_key=1, adx_key_seed=582, adx_key_mult=17765, adx_key_inc=28959, adx_encryption_type=8),
                    ct.convert_options(ct.CONTAINER_HCA, hca_quality=2, hca_key_type=56, hca_key_code=12345)):
            outs, status = ct.convert_wave_batch([f.numpy() for f in files], opt)
            print(opt.out_type, sum(o.size for o in outs), set(status))
        return 0
    if what == "c2":
        n_ch = 1024
        pcm = bench.make_batch_gpu(torch, n_ch, n, 0, dev)
        nb = vg.gcadpcm.sample_count_to_byte_count(n)
        a_stride = (nb + 15) // 16 * 16
        adpcm = torch.zeros((n_ch, a_stride), dtype=torch.uint8, device=dev)
        coefs = torch.zeros((n_ch, 16), dtype=torch.int16, device=dev)
        frames = (n + 13) // 14
        ws_bytes = int(vg.lib.vgb_gcadpcm_workspace_bytes(frames * n_ch, n_ch))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        off_p = np.arange(n_ch, dtype=np.int64) * n
        off_a = np.arange(n_ch, dtype=np.int64) * a_stride
        lens = np.full(n_ch, n, dtype=np.int32)
        N.check(vg.lib.vgb_gcadpcm_encode_dev(pcm.data_ptr(), off_p.ctypes.data, lens.ctypes.data, None, n_ch, None, coefs.data_ptr(),
                                              adpcm.data_ptr(), off_a.ctypes.data, ws.data_ptr(), ws_bytes, stream.cuda_stream))
    elif what == "gcdec":
        n_ch = 2048
        pcm = bench.make_batch_gpu(torch, n_ch, n, 1, dev).cpu().numpy()
        coefs, adpcm = vg.gcadpcm.encode_batch(pcm)
        dec = vg.gcadpcm.decode_batch(np.stack(adpcm), coefs, [vg.gcadpcm.GcAdpcmParameters(n)] * n_ch)
        vg.gcadpcm.seek_table_and_loop_context(adpcm[:256], coefs[:256], [n] * 256, 0x3800, [n // 3] * 256)
        del dec
    elif what == "frames":
        rng = np.random.default_rng(1)
        k = 65536
        io = rng.integers(-20000, 20000, (k, 16)).astype(np.int16)
        co = rng.integers(-2048, 2048, (k, 16)).astype(np.int16)
        out = np.zeros((k, 8), dtype=np.uint8)
        N.check(vg.lib.vgb_gcadpcm_encode_frames(io.ctypes.data, None, co.ctypes.data, k, out.ctypes.data))
    elif what == "adx":
        n_ch = 1024
        pcm = bench.make_batch_gpu(torch, n_ch, n, 2, dev).cpu().numpy()
        cfgs = [vg.criadx.CriAdxParameters()] * n_ch
        adpcm, hist = vg.criadx.encode_batch(pcm, cfgs)
        vg.criadx.decode_batch(adpcm, n, [vg.criadx.CriAdxParameters(history=int(h)) for h in hist])
    elif what == "hca":
        n_st = 128
        pcm = bench.make_batch_gpu(torch, n_st, n, 3, dev, degenerate=False).cpu().numpy()
        infos, frames = vg.crihca.encode_batch([[pcm[s]] for s in range(n_st)], RATE)
        vg.crihca.decode_batch(infos, frames)
        x = np.random.default_rng(2).standard_normal((64, 1024, 128))
        vg.crihca.mdct_run(vg.crihca.mdct_run(x), inverse=True)
    elif what == "ilv":
        items, count, size, ilv = 512, 2, 822864, 0x2000
        src = torch.randint(0, 256, (items, count, size), dtype=torch.uint8, device=dev)
        out = torch.zeros((items, count * size), dtype=torch.uint8, device=dev)
        back = torch.zeros((items, count, size), dtype=torch.uint8, device=dev)
        for tma in ("", "1"):
            if tma:
                os.environ["VGB_INTERLEAVE_TMA"] = "1"
            N.check(vg.lib.vgb_interleave_dev(src.data_ptr(), size, count * size, out.data_ptr(), count * size, items, count, size, ilv, size,
                                              stream.cuda_stream))
            N.check(vg.lib.vgb_deinterleave_dev(out.data_ptr(), count * size, back.data_ptr(), size, count * size, items, count, size, ilv, size,
                                                stream.cuda_stream))
    else:
        raise SystemExit(__doc__)
    torch.cuda.synchronize()
    print("done", what)


if __name__ == "__main__":
    main()
