"""Convert the models released with NanoCaller (TF2 checkpoints / Keras HDF5 under
<NanoCaller>/nanocaller_src/release_data) into this repo's packed `.ncw` format.

    python tools/convert_weights.py /path/to/NanoCaller/nanocaller_src

Output: nanocaller_b200/release_data/{snp,indel}/<model name>.ncw  (+ train_coverage in the header).
Runs wherever a NanoCaller checkout is available; the bundled copies were produced with it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanocaller_b200.host import weights as W  # noqa: E402


def main(src):
    for kind, table in (("snp", W.SNP_MODEL_DICT), ("indel", W.INDEL_MODEL_DICT)):
        outdir = os.path.join(W.RELEASE_DIR, kind)
        os.makedirs(outdir, exist_ok=True)
        for name, rel in table.items():
            if kind == "snp" and name == "NanoCaller2":
                continue   # alias of NanoCaller1 (snpCaller.py:17)
            path = os.path.join(src, rel)
            if rel.endswith(".h5"):
                tensors = W.read_keras_h5(path, kind)
                meta = {"train_coverage": 0.0, "haploid": True}
            else:
                tensors = W.read_tf_checkpoint(path)
                cov = path + ".coverage"
                meta = {"train_coverage": float(open(cov).readline().strip()) if os.path.exists(cov) else 0.0,
                        "haploid": False}
            meta.update(model=name, kind=kind, source=rel)
            out = W._ncw_path(kind, name)
            W.save_ncw(out, tensors, meta)
            print("%-6s %-28s -> %s (%d tensors, cov %.0f)" % (kind, name, os.path.relpath(out, ROOT), len(tensors), meta["train_coverage"]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/nanocaller_src")
