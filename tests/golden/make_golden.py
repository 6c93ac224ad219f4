"""Regenerates tests/golden/ from the reference checkout (run in the authoring container only;
/root/reference does not exist on the GPU box).

  mash_bounds_k27_p0.99.txt : stdout of `mash bounds -k 27 -p 0.99` as pinned by
                              /root/reference/tests/test_correct_workflow.py:26
  workflow_dicts.json       : the golden `mash_output.to_dict()` of the FASTQ and FASTA flows
                              (test_correct_workflow.py:18-25, :28-38) and the pinned scalars
"""
import ast
import json
import re
from pathlib import Path

REF = Path("/root/reference/tests/test_correct_workflow.py")
OUT = Path(__file__).resolve().parent
src = REF.read_text()
m = re.search(r'^error_bounds_text_ref = (".*")$', src, re.M)
(OUT / "mash_bounds_k27_p0.99.txt").write_text(ast.literal_eval(m.group(1)))
tree = ast.parse(src)
vals = {}
for node in tree.body:
    if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in (
            "mash_output_to_dict_fastq", "mash_output_to_dict_fasta"):
        vals[node.targets[0].id] = ast.literal_eval(node.value)
vals["fastq_estimated_genome_size"] = 48454.7      # :99
vals["fastq_minimal_distance"] = 9.55405e-06       # :105
vals["fasta_parsed_genome_size"] = 48502           # :197
vals["error_bound_s50000_d0.05"] = 0.0008979       # :129
(OUT / "workflow_dicts.json").write_text(json.dumps(vals, indent=1))
print("written", [p.name for p in OUT.iterdir()])
