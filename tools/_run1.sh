python -m pytest tests/test_crihca_gpu.py::test_mdct_taps_match_the_oracle_bit_for_bit -m gpu -q -x 2>&1 | grep -E "assert|Error|error" | head -8
prof() { # workload, kernel regex, launches, note, extra...
  w=$1; k=$2; c=$3; note=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$k" -c $c -f -o /tmp/r02_prof_$w python tools/profile_workloads.py $w > gpurun_out/r02_prof_$w.log 2>&1
  python tools/profile_kernels.py /tmp/r02_prof_$w.ncu-rep gpurun_out/profiles r02 "$note" "$@" 2>&1 | tail -8
}
prof c2 "gc_coef|gc_encode_kernel" 5 "C2: 1024 ch x 1440000 samples, coefficients + time-parallel encode (24 segments)" units=105326592 unit=frame
cp /tmp/r02_prof_c2.ncu-rep gpurun_out/r02_prof_c2.ncu-rep
prof gcdec "gc_decode_kernel" 2 "GC-ADPCM decode of 2048 ch x 1440000 samples; taps: 256 ch, seek table every 0x3800 samples + loop context" units=2949120000 unit=sample
prof frames "gc_encode_frames" 1 "DspEncodeFrame for 65536 independent frames" units=65536 unit=frame
prof adx "adx_" 4 "CRI ADX encode (time-parallel, 64 segments) + decode, 1024 ch x 1440000 samples" units=1474560000 unit=sample
prof hca "hca_" 6 "CRI HCA encode + decode of 128 mono streams x 1440000 samples (180096 frames); MDCT taps 64 x 1024 blocks" units=180096 unit=frame
prof ilv "interleave" 4 "block (de)interleave 512 x 2 x 822864 B, 0x2000-byte blocks; vector kernels then TMA bulk-copy kernels" units=1685225472 unit=byte
ls -la gpurun_out/profiles | head -40
