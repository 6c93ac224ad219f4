"""Issue rates of the instruction kinds kernel K2 is made of (thread-instructions per second, each kind issued
alone in independent chains) on the current GPU.  Run under gpurun; prints one JSON line."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from auriclass_b200.mash_gpu import MashGPU  # noqa: E402

g = MashGPU()
out = {"int_peak_imad_plus_lop3": g.int_peak(), **g.pipe_rates(), **g.alu_rates()}
print(json.dumps({k: round(v / 1e12, 3) for k, v in out.items()} | {"unit": "T thread-instr/s"}))
