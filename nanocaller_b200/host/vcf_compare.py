"""Record comparator implementing the tolerance rule of SURVEY.md D4: record keys, REF/ALT, FILTER, genotype and
all depth fields must be identical; QUAL and PR= derive from fp32 probabilities and are compared numerically
(PR within `tol`; QUAL within the change a `tol` shift of the probability can cause).  A record whose decision
hinges on a probability within `tol` of 0.5 is reported as `borderline` instead of `mismatch`."""
import math


def _parse(line):
    f = line.rstrip("\n").split("\t")
    info = dict(kv.split("=") for kv in f[7].split(";"))
    return {"key": (f[0], int(f[1])), "ref": f[3], "alt": f[4], "qual": float(f[5]), "filter": f[6],
            "pr": [float(x) for x in info["PR"].split(",")], "fq": info["FQ"], "fmt": f[8], "sample": f[9]}


def compare_records(got_lines, want_lines, tol=1e-4):
    """-> dict(n, identical, numeric_only, borderline, mismatch=[...])."""
    res = {"n": len(want_lines), "identical": 0, "numeric_only": 0, "borderline": 0, "mismatch": []}
    if len(got_lines) != len(want_lines):
        res["mismatch"].append("record count %d != %d" % (len(got_lines), len(want_lines)))
        return res
    for g, w in zip(got_lines, want_lines):
        if g == w:
            res["identical"] += 1
            continue
        a, b = _parse(g), _parse(w)
        pr_ok = all(abs(x - y) <= tol + 1e-4 + 1e-9 for x, y in zip(a["pr"], b["pr"]))     # printed with 4 decimals: two values on either
                                                                                          # side of a rounding boundary print one unit apart
        same_call = all(a[k] == b[k] for k in ("key", "ref", "alt", "filter", "fq", "fmt", "sample"))
        if same_call and pr_ok:
            # QUAL = -10 log10(1e-10 + 1 - p): d(QUAL)/dp = 10 / (ln 10 (1 - p)); allow the change caused by a tol shift
            slack = 0.002 + 10.0 / math.log(10) * tol / max(1e-10, 10 ** (-max(a["qual"], b["qual"]) / 10.0))
            if abs(a["qual"] - b["qual"]) <= slack or a["filter"] == "LOW":
                res["numeric_only"] += 1
                continue
        if a["key"] == b["key"] and pr_ok and any(abs(p - 0.5) <= tol for p in b["pr"]):
            res["borderline"] += 1
            continue
        res["mismatch"].append((g, w))
    return res
