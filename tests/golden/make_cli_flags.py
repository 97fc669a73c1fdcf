"""Extract the flag names and defaults of the reference's command line (/root/reference/NanoCaller, argparse add_argument calls)
into tests/golden/reference_cli_flags.txt (one `--flag<TAB>default repr` per line).  Build container only (the GPU box has no
/root/reference)."""
import ast
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/NanoCaller").read()
rows = {}
for m in re.finditer(r"add_argument\((.*)\)\s*$", src, flags=re.M):
    call = m.group(1)
    names = re.findall(r"['\"](--[A-Za-z0-9_]+)['\"]", call)
    if not names:
        continue
    d = re.search(r"default\s*=\s*('[^']*'|\"[^\"]*\"|[^,\)]+)", call)
    rows[names[0]] = repr(ast.literal_eval(d.group(1).strip())) if d else "None"
with open(os.path.join(HERE, "reference_cli_flags.txt"), "w") as f:
    for k in sorted(rows):
        f.write("%s\t%s\n" % (k, rows[k]))
print(len(rows), "flags")
