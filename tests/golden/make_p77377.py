"""Regenerates tests/golden/p77377_trypsin.json from the reference's own test file
(src/proteomic/models/enzyms/tests/digest_enzym.rs:13-92).  Run in the build container only
(/root/reference does not exist on the GPU box)."""
import json
import re

src = open('/root/reference/src/proteomic/models/enzyms/tests/digest_enzym.rs').read()
seq = re.search(r'P77377_SEQUENCE: &str = "([A-Z]+)"', src).group(1)
peps = re.findall(r'^\s+"([A-Z]+)",?$', src, flags=re.M)
assert len(peps) == 71
json.dump({"source": "src/proteomic/models/enzyms/tests/digest_enzym.rs:13-92", "protein": "P77377", "sequence": seq,
           "params": {"max_missed_cleavages": 2, "min_len": 6, "max_len": 50}, "peptides": peps},
          open(__file__.replace('make_p77377.py', 'p77377_trypsin.json'), 'w'), indent=1)
