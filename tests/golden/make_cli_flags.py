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

# the preset table (NanoCaller:66-77) as JSON
import json

i = src.index("preset_dict={") + len("preset_dict=")
depth = 0
for j in range(i, len(src)):
    if src[j] == "{":
        depth += 1
    elif src[j] == "}":
        depth -= 1
        if depth == 0:
            break
json.dump(ast.literal_eval(src[i:j + 1]), open(os.path.join(HERE, "reference_presets.json"), "w"), indent=1, sort_keys=True)

# VCF header lines written by snpCaller.call_manager (snpCaller.py:259-276) and indelCaller.call_manager (indelCaller.py:373-383)
hdr = {}
for name, fn in (("snps", "snpCaller.py"), ("indels", "indelCaller.py")):
    text = open("/root/reference/nanocaller_src/" + fn).read()
    lines = re.findall(r"outfile\.write\(b'((?:##|#CHROM)[^']*)'", text)
    hdr[name] = [ln.replace("\\n", "").replace("\\t", "\t") for ln in lines]
json.dump(hdr, open(os.path.join(HERE, "reference_vcf_headers.json"), "w"), indent=1)

# model name tables (snpCaller.py:16-34, indelCaller.py:17-24)
tables = {}
for key, fn, var in (("snp", "snpCaller.py", "snp_model_dict"), ("indel", "indelCaller.py", "indel_model_dict")):
    text = open("/root/reference/nanocaller_src/" + fn).read()
    i = text.index(var + "={") + len(var) + 1
    tables[key] = ast.literal_eval(text[i:text.index("}", i) + 1])
json.dump(tables, open(os.path.join(HERE, "reference_model_tables.json"), "w"), indent=1, sort_keys=True)
