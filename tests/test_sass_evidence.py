"""Build-time evidence for the design claims of DESIGN.md section 5, read from the compiled objects (no GPU needed):
K2 stages its sub-tile with a TMA bulk copy, its flat hot loop is still ONE basic block (a compiler or source change
that re-introduces a branch per window would show here long before it shows in a bench), and nothing spills."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "auriclass_b200" / "csrc" / "ac_hash.o"
LOG = ROOT / "auriclass_b200" / "csrc" / "ac_hash.ptxas.log"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="CUDA toolkit not on PATH")


@pytest.fixture(scope="module")
def k2_sass():
    if not OBJ.exists():
        import __graft_entry__ as g
        g.build()
    out = subprocess.run(["cuobjdump", "-sass", str(OBJ)], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and name:
            funcs[name].append(m.group(1).strip())
    return funcs


def _kernel(funcs, k, flat):
    hits = [v for n, v in funcs.items() if f"k2_hashILi{k}ELb{1 if flat else 0}E" in n]
    assert len(hits) == 1, f"k2_hash<{k}, {flat}> not found in ac_hash.o"
    return hits[0]


def _longest_straight_block(ins):
    best = cur = 0
    for t in ins:
        if re.search(r"\b(BRA|CALL|BSSY|BSYNC|EXIT|RET|WARPSYNC)\b", t):
            best, cur = max(best, cur), 0
        else:
            cur += 1
    return max(best, cur)


@pytest.mark.parametrize("k", [16, 21, 27, 32])
def test_k2_uses_tma_and_flat_loop_is_one_block(k2_sass, k):
    flat, branchy = _kernel(k2_sass, k, True), _kernel(k2_sass, k, False)
    for ins in (flat, branchy):
        assert any("UBLKCP" in t for t in ins), "no TMA bulk copy (UBLKCP)"
        assert any("SYNCS" in t for t in ins), "no mbarrier"
    # 8 windows x (14 IMAD.WIDE + ...) in one block; the branchy loop has a branch and a call per window
    n_flat, n_branchy = _longest_straight_block(flat), _longest_straight_block(branchy)
    assert n_flat >= 600, n_flat
    assert n_branchy <= 320 and n_branchy < n_flat // 2, (n_branchy, n_flat)   # (its longest block is the unrolled RC / bitmap prologue)
    wide = sum(1 for t in flat if re.match(r"(@!?U?P\d\s+)?IMAD\.WIDE", t))
    assert wide >= 8 * (14 if k > 16 else 10), wide   # the 64-bit multiplies of the 8 unrolled windows


def test_k2_does_not_spill_at_the_production_k():
    if not LOG.exists():
        pytest.skip("no ptxas log (library built elsewhere)")
    txt = LOG.read_text()
    blocks = re.split(r"ptxas info\s+: Compiling entry function ", txt)
    seen = 0
    for b in blocks:
        if re.match(r"'\S*k2_hashILi27ELb[01]E", b):
            seen += 1
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
            assert m and m.group(2) == "0" and m.group(3) == "0", b[:300]
            regs = int(re.search(r"Used (\d+) registers", b).group(1))
            assert regs <= 64, regs   # 8 CTAs of 128 threads per SM
    assert seen == 2


def test_k1_fastq_parser_stages_its_tile_with_tma_and_fits_five_ctas():
    """north_star (a): the record parser's text goes through TMA / shared memory -- k1q_fused brings its 16 KiB tile in
    with one bulk copy on an mbarrier; <= 51 registers x 256 threads and ~37 KB of shared memory keep 5 CTAs on an SM."""
    obj = ROOT / "auriclass_b200" / "csrc" / "ac_parse.o"
    if not obj.exists():
        import __graft_entry__ as g
        g.build()
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    body, on = [], False
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            on = "k1q_fused" in m.group(1)
        elif on:
            body.append(line)
    assert body, "k1q_fused not found in ac_parse.o"
    txt = "\n".join(body)
    assert "UBLKCP" in txt and "SYNCS" in txt
    assert "REDUX" in txt          # the look-back folds 32 predecessors with warp reductions
    log = ROOT / "auriclass_b200" / "csrc" / "ac_parse.ptxas.log"
    if log.exists():
        for b in re.split(r"ptxas info\s+: Compiling entry function ", log.read_text()):
            if "k1q_fused" in b[:200]:
                regs = int(re.search(r"Used (\d+) registers", b).group(1))
                smem = int(re.search(r"(\d+) bytes smem", b).group(1))
                assert regs <= 51 and smem <= 40 * 1024, (regs, smem)


def test_k5_pairs_prefetches_with_ldgsts_and_walks_without_branches():
    """DESIGN.md section 5 (K5, all-vs-all): the next block is requested with asynchronous global -> shared copies
    (LDGSTS) while the warp walks, and the walk of a full block is straight-line code -- P_PER unrolled steps of one
    shared-memory load, two 64-bit compares and selects, no branch (the divergent walk of round 1 ran at 15 of 32 lanes)."""
    obj = ROOT / "auriclass_b200" / "csrc" / "ac_dist.o"
    if not obj.exists():
        import __graft_entry__ as g
        g.build()
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    ins, on = [], False
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            on = "k5_pairs" in m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and on:
            ins.append(m.group(1).strip())
    assert ins, "k5_pairs not found in ac_dist.o"
    assert sum(1 for t in ins if "LDGSTS" in t) >= 2 * 13, "staging is not asynchronous"
    assert any("LDGDEPBAR" in t or "DEPBAR" in t for t in ins), "no wait on the copy group"
    # the longest straight block holds the unrolled walk (its first step or two are scheduled into the code that picks the
    # "A value before"): >= 11 of the 13 steps x (load + 4 compare instructions + selects)
    best, cur, loads, best_loads = 0, 0, 0, 0
    for t in ins:
        if re.search(r"\b(BRA|CALL|BSSY|BSYNC|EXIT|RET|WARPSYNC)\b", t):
            if cur > best:
                best, best_loads = cur, loads
            cur = loads = 0
        else:
            cur += 1
            loads += 1 if re.match(r"(@!?P\d\s+)?LDS\.64", t) else 0
    if cur > best:
        best, best_loads = cur, loads
    assert best >= 11 * 13 and best_loads >= 11, (best, best_loads)
    log = ROOT / "auriclass_b200" / "csrc" / "ac_dist.ptxas.log"
    if log.exists():
        for b in re.split(r"ptxas info\s+: Compiling entry function ", log.read_text()):
            if "k5_pairs" in b[:200]:
                m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
                assert m and m.group(2) == "0" and m.group(3) == "0", b[:300]
                assert int(re.search(r"Used (\d+) registers", b).group(1)) <= 64
