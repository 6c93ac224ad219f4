python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dist or pairs or batch" > gpurun_out/k5_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/k5_pytest.log
for v in w4_p17 w8_p9 w4_p9 w2_p33 w4_p13 w8_p5; do
  AURICLASS_CUDA_LIB=$PWD/gpurun_variants/$v.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k5_pairs --csv --log-file gpurun_out/k5_$v.csv python tools/bench_batch.py --skip c4 --sketches 768 --tile 256 --check 2 --out gpurun_out/k5_$v.jsonl > /dev/null 2>&1
  python - <<PY
import csv,json
r=[x for x in csv.reader(open('gpurun_out/k5_$v.csv')) if len(x)>5 and 'k5_pairs' in ' '.join(x)]
t=[float(x[-1].replace(',','')) for x in r]
big=[x for x in t if x>max(t)*0.5]
d=json.loads(open('gpurun_out/k5_$v.jsonl').readline())
print('$v kernel', len(t), 'launches; full tiles:', [round(x/1e6,2) for x in big], 'ms -> M pairs/s', round(65536/(sum(big)/len(big))*1e3,2), 'parity', d['parity']['equal'])
PY
done
ncu --set full --clock-control none --import-source on -k regex:k5_pairs -s 2 -c 1 -o gpurun_out/k5_pairs_new3 -f python tools/bench_batch.py --skip c4 --sketches 512 --tile 256 --check 0 > gpurun_out/k5_ncu.log 2>&1; echo ncu rc=$?
