"""Extract the flag names of the reference's command line (/root/reference/NanoCaller, argparse add_argument calls) into
tests/golden/reference_cli_flags.txt.  Build container only (the GPU box has no /root/reference)."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/NanoCaller").read()
flags = sorted(set(re.findall(r"add_argument\(\s*['\"](--?[A-Za-z0-9_\-]+)['\"]", src)))
open(os.path.join(HERE, "reference_cli_flags.txt"), "w").write("\n".join(flags) + "\n")
print(len(flags), "flags")
