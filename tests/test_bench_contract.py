"""The driver parses ONE JSON line from bench.py.  The reference arm runs on the CPU, so its line is checked here;
the keys of our own arm's line are checked against the source (it needs a GPU to run)."""
import json
import re
import sys
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-sample-reads", "20000", "--n-refs", "2"], capture_output=True, text=True, cwd=str(ROOT),
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "FASTQ Gbases/s sketched (-m filter)" and d["unit"] == "Gbases/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    # both arms name the workload with the SAME string (the driver compares them)
    import bench
    assert d["config"]["workload"] == bench.WORKLOADS["c2"]
    src = (ROOT / "bench.py").read_text()
    assert src.count('"workload": WORKLOADS[args.config]') == 2


def test_our_arm_emits_every_contract_key_and_matches_the_library_defaults():
    src = (ROOT / "bench.py").read_text()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"e2e"', '"h2d_bytes_per_step"',
                '"d2h_bytes_per_step"', '"gpu_launches"', '"clocks"', '"roofline"', '"cpu_baseline"', '"traffic"', '"frac"', '"parity"',
                '"equal"', '"stages_ms"', '"inflate"'):
        assert key in src, key
    # bench.py derives K2 launches per step from the segment size: keep it in step with the library's default
    seg = re.search(r"constexpr size_t SEG_BYTES = (\d+)u << 20;", (ROOT / "auriclass_b200/csrc/ac_common.cuh").read_text())
    assert seg and f'os.environ.get("AC_SEG_MB", "{seg.group(1)}")' in src
