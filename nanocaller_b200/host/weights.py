"""Model-weight loading for the four NanoCaller CNNs without TensorFlow or h5py.

Readers for the two on-disk formats the reference ships under nanocaller_src/release_data
(snpCaller.py:16-34 `snp_model_dict`, indelCaller.py:17-24 `indel_model_dict`):

  * TF2 object checkpoints (`<prefix>.index` LevelDB table + `<prefix>.data-00000-of-00001`);
  * Keras HDF5 files (superblock v0, v1 object headers, contiguous float32 datasets).

plus this repo's own packed format `.ncw` (a flat little-endian float32 blob with a JSON header)
in which converted copies of the released models are stored under nanocaller_b200/release_data,
so the package is usable without a NanoCaller checkout.  Layer naming is normalised to the
attribute names of model_architect*.py: conv1_1 conv1_2 conv1_3 conv2 conv3 fc1 fa A G T C fc2 fc3 GT.
"""
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RELEASE_DIR = os.path.join(os.path.dirname(HERE), "release_data")

# model name -> path relative to the reference's nanocaller_src/ (snpCaller.py:16-34)
SNP_MODEL_DICT = {
    "NanoCaller1": "release_data/ONT_models/SNPs/NanoCaller1_beta/model-rt-1",
    "NanoCaller2": "release_data/ONT_models/SNPs/NanoCaller1_beta/model-rt-1",   # sic: reference maps 2 -> 1
    "NanoCaller3": "release_data/clr_models/SNPs/NanoCaller3_beta/model-rt-100",
    "ONT-HG001": "release_data/ONT_models/SNPs/HG001_guppy4.2.2_giab-3.3.2/model-1",
    "ONT-HG001_GP2.3.8": "release_data/ONT_models/SNPs/HG001_guppy2.3.8_giab-3.3.2/model-100",
    "ONT-HG001_GP2.3.8-4.2.2": "release_data/ONT_models/SNPs/HG001_guppy2.3.8_guppy4.2.2_giab-3.3.2/model-100",
    "ONT-HG001-4_GP4.2.2": "release_data/ONT_models/SNPs/HG001_guppy4.2.2_giab-3.3.2_HG002-4_guppy4.2.2_giab-4.2.1/model-100",
    "ONT-HG002": "release_data/ONT_models/SNPs/HG002_guppy4.2.2_giab-4.2.1/model-100",
    "ONT-HG002_GP4.2.2_v3.3.2": "release_data/ONT_models/SNPs/HG002_guppy4.2.2_giab-3.3.2/model-100",
    "ONT-HG002_GP2.3.4_v3.3.2": "release_data/ONT_models/SNPs/HG002_guppy2.3.4_giab-3.3.2/model-100",
    "ONT-HG002_GP2.3.4_v4.2.1": "release_data/ONT_models/SNPs/HG002_guppy2.3.4_giab-4.2.1/model-100",
    "ONT-HG002_r10.3": "release_data/ONT_models/SNPs/HG002_r10.3_guppy4.0.11_giab-4.2.1/model-100",
    "ONT-HG002_bonito": "release_data/ONT_models/SNPs/HG002_bonito_giab-4.2.1/model-100",
    "CCS-HG001": "release_data/hifi_models/SNPs/HG001_giab-3.3.2/model-100",
    "CCS-HG002": "release_data/hifi_models/SNPs/HG002_giab-4.2.1/model-100",
    "CCS-HG001-4": "release_data/hifi_models/SNPs/HG001_giab-3.3.2_HG002-4_giab-4.2.1/model-100",
    "CLR-HG002": "release_data/clr_models/SNPs/HG002_giab-4.2.1/model-100",
    "haploid": "release_data/haploid_models/SNPs/CHM13/model.24-0.9985.h5",
}
# indelCaller.py:17-24
INDEL_MODEL_DICT = {
    "NanoCaller1": "release_data/ONT_models/indels/NanoCaller1_beta/model-30",
    "NanoCaller3": "release_data/hifi_models/indels/NanoCaller3_beta/model-25",
    "ONT-HG001": "release_data/ONT_models/indels/HG001_guppy4.2_giab-3.3.2/model-100",
    "ONT-HG002": "release_data/ONT_models/indels/HG002_guppy4.2_giab-4.2.1/model-100",
    "CCS-HG001": "release_data/hifi_models/indels/HG001_giab-3.3.2/model-100",
    "CCS-HG002": "release_data/hifi_models/indels/HG002_giab-4.2.1/model-100",
    "haploid": "release_data/haploid_models/indels/CHM13/model.19-0.9811.h5",
}

SNP_LAYERS = ["conv1_1", "conv1_2", "conv1_3", "conv2", "conv3", "fc1", "fa", "A", "G", "T", "C", "fc2", "fc3", "GT"]
SNP_HAP_LAYERS = ["conv1_1", "conv1_2", "conv1_3", "conv2", "conv3", "fc1", "fc2", "fc3"]
INDEL_LAYERS = ["conv1_1", "conv1_2", "conv1_3", "conv2", "conv3", "fc1", "fc2", "fc3"]
# Keras layer names inside the .h5 files -> attribute names (loaded by layer order in the reference)
_H5_SNP = {"C1_1": "conv1_1", "C1_2": "conv1_2", "C1_3": "conv1_3", "C2": "conv2", "C3": "conv3",
           "C4": "fc1", "C6": "fc2", "C7": "fc3"}
_H5_INDEL = {"C1_1": "conv1_1", "C1_2": "conv1_2", "C1_3": "conv1_3", "C2": "conv2", "C3": "conv3",
             "C4": "fc1", "C5": "fc2", "C6": "fc3"}


# ------------------------------------------------------------------ TF2 checkpoint bundle
def _varint(buf, i):
    shift = val = 0
    while True:
        b = buf[i]
        i += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, i
        shift += 7


def _block_entries(buf, off, size):
    """Entries of one LevelDB table block (no compression): prefix-compressed keys."""
    block = buf[off:off + size]
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    i, key = 0, b""
    while i < end:
        shared, i = _varint(block, i)
        non_shared, i = _varint(block, i)
        vlen, i = _varint(block, i)
        key = key[:shared] + bytes(block[i:i + non_shared])
        i += non_shared
        yield key, bytes(block[i:i + vlen])
        i += vlen


def _parse_bundle_entry(val):
    """BundleEntryProto: 1 dtype, 2 shape{2 dim{1 size}}, 3 shard, 4 offset, 5 size, 6 crc32c."""
    out = {"dtype": 0, "shape": [], "offset": 0, "size": 0}
    i = 0
    while i < len(val):
        tag, i = _varint(val, i)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, i = _varint(val, i)
            if field == 1:
                out["dtype"] = v
            elif field == 4:
                out["offset"] = v
            elif field == 5:
                out["size"] = v
        elif wt == 2:
            ln, i = _varint(val, i)
            sub = val[i:i + ln]
            i += ln
            if field == 2:
                j = 0
                while j < len(sub):
                    t2, j = _varint(sub, j)
                    if t2 & 7 == 2:
                        l2, j = _varint(sub, j)
                        dim = sub[j:j + l2]
                        j += l2
                        if t2 >> 3 == 2:
                            k, size = 0, 0
                            while k < len(dim):
                                t3, k = _varint(dim, k)
                                if t3 & 7 == 0:
                                    v3, k = _varint(dim, k)
                                    if t3 >> 3 == 1:
                                        size = v3
                                elif t3 & 7 == 2:
                                    l3, k = _varint(dim, k)
                                    k += l3
                            out["shape"].append(size)
                    elif t2 & 7 == 0:
                        _, j = _varint(sub, j)
        elif wt == 5:
            i += 4
        elif wt == 1:
            i += 8
    return out


def read_tf_checkpoint(prefix):
    """-> {'<attr>/kernel': ndarray, '<attr>/bias': ndarray, ...} from `<prefix>.index` + data shard."""
    idx = open(prefix + ".index", "rb").read()
    if idx[-8:] != bytes.fromhex("57fb808b247547db"):
        raise ValueError("%s.index: not a TF checkpoint index (bad table magic)" % prefix)
    footer = idx[-48:]
    i = 0
    _, i = _varint(footer, i)
    _, i = _varint(footer, i)
    index_off, i = _varint(footer, i)
    index_size, i = _varint(footer, i)
    data = np.fromfile(prefix + ".data-00000-of-00001", dtype=np.uint8)
    tensors = {}
    for _, handle in _block_entries(idx, index_off, index_size):
        boff, j = _varint(handle, 0)
        bsize, j = _varint(handle, j)
        for key, val in _block_entries(idx, boff, bsize):
            if not key.endswith(b"/.ATTRIBUTES/VARIABLE_VALUE"):
                continue
            e = _parse_bundle_entry(val)
            if e["dtype"] != 1:
                continue
            name = key[:-len(b"/.ATTRIBUTES/VARIABLE_VALUE")].decode()
            arr = data[e["offset"]:e["offset"] + e["size"]].view("<f4").reshape(e["shape"]).copy()
            tensors[name] = arr
    return tensors


# ------------------------------------------------------------------ Keras HDF5 (superblock v0)
class _H5:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        if self.b[:8] != b"\x89HDF\r\n\x1a\n" or self.b[8] != 0:
            raise ValueError("%s: only HDF5 superblock version 0 is supported" % path)
        # root group symbol-table entry starts at byte 56: link name off(8), object header addr(8)
        self.root = struct.unpack_from("<Q", self.b, 56 + 8)[0]

    def _messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg + 64:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end:
                mtype, msize, _ = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((coff, clen))
                else:
                    out.append((mtype, body, msize))
                p = body + msize
        return out

    def _group_entries(self, addr):
        b = self.b
        for mtype, body, _ in self._messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", b, body)
                heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
                return dict(self._walk_btree(btree, heap_data))
        return {}

    def _walk_btree(self, addr, heap_data):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        level = b[addr + 5]
        n = struct.unpack_from("<H", b, addr + 6)[0]
        for i in range(n):
            child = struct.unpack_from("<Q", b, addr + 24 + 8 + 16 * i)[0]
            if level > 0:
                yield from self._walk_btree(child, heap_data)
            else:
                assert b[child:child + 4] == b"SNOD"
                cnt = struct.unpack_from("<H", b, child + 6)[0]
                for k in range(cnt):
                    e = child + 8 + 40 * k
                    name_off, ohdr = struct.unpack_from("<QQ", b, e)
                    s = heap_data + name_off
                    name = b[s:b.index(b"\0", s)].decode()
                    yield name, ohdr

    def _dataset(self, addr):
        b = self.b
        shape, daddr, dsize = None, None, None
        for mtype, body, _ in self._messages(addr):
            if mtype == 0x01:
                ver, rank = b[body], b[body + 1]
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, b, off)
            elif mtype == 0x03:
                cls = b[body] & 0x0F
                size = struct.unpack_from("<I", b, body + 4)[0]
                if cls != 1 or size != 4:
                    return None
            elif mtype == 0x08:
                ver, lclass = b[body], b[body + 1]
                if ver == 3 and lclass == 1:
                    daddr, dsize = struct.unpack_from("<QQ", b, body + 2)
        if shape is None or daddr is None:
            return None
        return np.frombuffer(b, "<f4", count=dsize // 4, offset=daddr).reshape(shape).copy()

    def walk(self, addr=None, prefix=""):
        addr = self.root if addr is None else addr
        for name, ohdr in self._group_entries(addr).items():
            path = prefix + "/" + name
            sub = self._group_entries(ohdr)
            if sub:
                yield from self.walk(ohdr, path)
            else:
                arr = self._dataset(ohdr)
                if arr is not None:
                    yield path, arr


def read_keras_h5(path, kind):
    """kind 'snp' | 'indel' -> {'<attr>/kernel': ..., '<attr>/bias': ...}."""
    table = _H5_SNP if kind == "snp" else _H5_INDEL
    out = {}
    for p, arr in _H5(path).walk():
        parts = p.strip("/").split("/")
        layer, leaf = parts[0], parts[-1]
        if layer in table and leaf in ("kernel:0", "bias:0"):
            out["%s/%s" % (table[layer], leaf[:-2])] = arr
    return out


# ------------------------------------------------------------------ .ncw packed format
def save_ncw(path, tensors, meta):
    names = sorted(tensors)
    header = {"meta": meta, "tensors": []}
    off = 0
    for n in names:
        a = np.ascontiguousarray(tensors[n], dtype="<f4")
        header["tensors"].append({"name": n, "shape": list(a.shape), "offset": off})
        off += a.size
    hb = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(b"NCW1" + struct.pack("<I", len(hb)) + hb)
        f.write(b"\0" * ((-(8 + len(hb))) % 16))
        for n in names:
            f.write(np.ascontiguousarray(tensors[n], dtype="<f4").tobytes())


def load_ncw(path):
    raw = open(path, "rb").read()
    if raw[:4] != b"NCW1":
        raise ValueError("%s: not an .ncw weight file" % path)
    hl = struct.unpack_from("<I", raw, 4)[0]
    header = json.loads(raw[8:8 + hl])
    base = 8 + hl + ((-(8 + hl)) % 16)
    blob = np.frombuffer(raw, "<f4", offset=base)
    tensors = {}
    for t in header["tensors"]:
        n = int(np.prod(t["shape"])) if t["shape"] else 1
        tensors[t["name"]] = blob[t["offset"]:t["offset"] + n].reshape(t["shape"]).copy()
    return tensors, header["meta"]


# ------------------------------------------------------------------ model lookup (get_SNP_model / get_indel_model)
def _ncw_path(kind, name):
    if kind == "snp" and name == "NanoCaller2":   # snpCaller.py:17 maps NanoCaller2 onto the NanoCaller1 weights
        name = "NanoCaller1"
    return os.path.join(RELEASE_DIR, kind, name.replace("/", "_") + ".ncw")


def load_model(kind, name, nanocaller_src=None):
    """kind 'snp' | 'indel'.  `name` is a key of the reference's model dicts or a path to a model
    directory / checkpoint prefix / .h5 / .ncw.  Returns (tensors, meta) where meta has
    'train_coverage' (float, 0 when no .coverage sidecar — snpCaller.py:48-53) and 'haploid' (bool).
    Mirrors snpCaller.get_SNP_model (snpCaller.py:36-55) / indelCaller.get_indel_model (:26-38);
    returns (None, None) for an unknown name like the reference does."""
    table = SNP_MODEL_DICT if kind == "snp" else INDEL_MODEL_DICT
    if name in table:
        p = _ncw_path(kind, name)
        if os.path.exists(p):
            return load_ncw(p)
        if nanocaller_src is None:
            raise FileNotFoundError("model %r is not bundled (%s) and no NanoCaller checkout was given" % (name, p))
        path = os.path.join(nanocaller_src, table[name])
    elif os.path.exists(name) or os.path.exists(name + ".index"):
        path = name
        if os.path.isdir(path):
            import glob
            hits = sorted(glob.glob(os.path.join(path, "*.index")))
            if not hits:
                return None, None
            path = hits[0][:-len(".index")]
    else:
        return None, None
    if path.endswith(".ncw"):
        return load_ncw(path)
    if path.endswith(".h5"):
        return read_keras_h5(path, kind), {"train_coverage": 0.0, "haploid": True, "source": os.path.basename(path)}
    tensors = read_tf_checkpoint(path)
    cov_path = path + ".coverage"
    cov = float(open(cov_path).readlines()[0].rstrip("\n")) if os.path.exists(cov_path) else 0.0
    return tensors, {"train_coverage": cov, "haploid": False, "source": os.path.basename(path)}


# ------------------------------------------------------------------ packed blobs for the C-ABI
def _pack(tensors, layers):
    parts = []
    for name in layers:
        parts.append(np.ascontiguousarray(tensors[name + "/kernel"], dtype="<f4").ravel())
        parts.append(np.ascontiguousarray(tensors[name + "/bias"], dtype="<f4").ravel())
    return np.concatenate(parts)


def pack_snp_blob(tensors, haploid):
    """Canonical tensor order of nc_load_snp_weights (include/nanocaller_b200.h)."""
    return _pack(tensors, SNP_HAP_LAYERS if haploid else SNP_LAYERS)


def pack_indel_blob(tensors):
    """Canonical tensor order of nc_load_indel_weights."""
    return _pack(tensors, INDEL_LAYERS)
