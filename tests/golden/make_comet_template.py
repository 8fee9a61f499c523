"""Extracts the two constant blocks of the Comet parameter file the reference emits
(/root/reference/src/proteomic/utility/comet_parameter.rs: COMET_PARAMS_BEGIN / COMET_PARAMS_END) into a fixture, so the
byte-compatibility test of maxdecoy.outputs.comet_params can run where the reference tree is absent.
Run in the build container:  python tests/golden/make_comet_template.py"""
import json
import os
import re

SRC = "/root/reference/src/proteomic/utility/comet_parameter.rs"
text = open(SRC).read()
begin = re.search(r'const COMET_PARAMS_BEGIN: &\'static str = "(.*?)";', text, re.S).group(1)
end = re.search(r'const COMET_PARAMS_END: &\'static str = "(.*?)";', text, re.S).group(1)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "comet_params_template.json")
with open(out, "w") as fh:
    json.dump({"source": "src/proteomic/utility/comet_parameter.rs:6-94", "begin": begin, "end": end}, fh, indent=1)
print("wrote", out, len(begin), len(end))
