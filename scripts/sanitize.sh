#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the wavefront kernel's hardest parity tests and one dense /
# one generic-sparse test. Run on the GPU box: bash scripts/sanitize.sh [outdir]. Logs go to gpurun_out/sanitizer/.
OUT=${1:-gpurun_out/sanitizer}
mkdir -p "$OUT"
SEL='test_sparse_wavefront_stress or test_sparse_wavefront_dense_conflicts_and_repeated_samples'
SEL2='test_dense_multinomial_wine or test_sparse_generic_multiclass or test_sparse_long_and_empty_rows'
for tool in memcheck racecheck synccheck; do
  for grp in wave other; do
    if [ $grp = wave ]; then K="$SEL"; else K="$SEL2"; fi
    timeout 1500 compute-sanitizer --tool $tool --print-limit 50 --log-file "$OUT/${tool}_${grp}.log" \
      python -m pytest tests/test_parity_gpu.py -x -q -k "$K" > "$OUT/${tool}_${grp}.pytest.txt" 2>&1
    echo "$tool $grp rc=$? $(tail -n 1 "$OUT/${tool}_${grp}.log")" | tee -a "$OUT/summary.txt"
  done
done
