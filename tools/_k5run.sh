python -m pytest tests/test_gpu_parity.py tests/test_reference_workflow.py -m gpu -q -x -k "dist or pairs or batch or report or reference" > gpurun_out/k5_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/k5_pytest.log
rm -f gpurun_out/batch_r2s_n1.jsonl
python tools/bench_batch.py --assemblies 96 --sketches 1024 --tile 256 --out gpurun_out/batch_r2s_n1.jsonl > /dev/null 2> gpurun_out/batch_r2s.err; echo batch rc=$?
python -c "
import json
for l in open('gpurun_out/batch_r2s_n1.jsonl'):
    d=json.loads(l); print(d['path'], d.get('assemblies_per_s'), d.get('pairs_per_s'), d.get('tile'), d['parity']['equal'], d['rank0_calls'], d.get('device_pairs_per_s_rank0'))"
