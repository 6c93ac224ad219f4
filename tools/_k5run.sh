rm -f gpurun_out/batch_r2s_n1.jsonl
python tools/bench_batch.py --assemblies 96 --sketches 1024 --tile 256 --out gpurun_out/batch_r2s_n1.jsonl > /dev/null 2> gpurun_out/batch_r2s.err; echo batch rc=$?
python -c "
import json
for l in open('gpurun_out/batch_r2s_n1.jsonl'):
    d=json.loads(l); print(d['path'], d.get('assemblies_per_s'), d.get('pairs_per_s'), d.get('tile'), d['parity']['equal'], d['rank0_calls'], d.get('device_pairs_per_s_rank0'))"
