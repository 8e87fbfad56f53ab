# sample-size experiment (QSAE_SAMPLE_DIV): stage timings and the two-stream step
for div in 12 8; do
  for k in 32 65; do
    echo "== div=$div k=$k"; QSAE_SAMPLE_DIV=$div timeout 120 python tools/prof_stages.py 4096 $k 40 | tail -6
    QSAE_SAMPLE_DIV=$div timeout 120 python tools/prof_overlap.py 4096 $k 200 2>/dev/null | head -2
  done
done
for div in 32 16; do
  for k in 32 65; do
    echo "== B=65536 div=$div k=$k"; QSAE_SAMPLE_DIV=$div timeout 120 python tools/prof_stages.py 65536 $k 10 | tail -7
  done
  echo "== B=65536 div=$div k=32 exact"; QSAE_SAMPLE_DIV=$div timeout 120 python tools/prof_stages.py 65536 32 10 1 | tail -7
done
