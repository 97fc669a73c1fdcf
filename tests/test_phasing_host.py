"""CPU: read-based phasing + haplotagging (libnc_phase.so, stands in for `whatshap phase` / `haplotag`, indelCaller.py:237,:244)
against the synthetic generator's true haplotypes: the generator tags every read with the haplotype it was drawn from."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world(**kw):
    from nanocaller_b200.synth import make_world
    return make_world(**kw)


def _truth_lines(w, qual=50.0, hom_too=True):
    pos1, kinds, alts = w.truth_snps()
    lines = []
    for p, k, a in zip(pos1.tolist(), kinds.tolist(), alts.tolist()):
        ref = chr(w.reads.ref[p - 1])
        if ref not in "ACGT" or (k == 2 and not hom_too):
            continue
        gt = "0/1" if k == 1 else "1/1"
        lines.append("%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:DP:VF:AD:ADF:ADR\t%s:30:0.5000:15,15:8,7:7,8\n" % (w.reads.chrom, p, ref, chr(a), qual, gt))
    return lines


def _block_agreement(truth_hp, got_hp, got_ps):
    """Reads tagged by both: agreement with the truth up to a swap per phase block."""
    both = (truth_hp > 0) & (got_hp > 0)
    agree = total = 0
    for ps in np.unique(got_ps[both]):
        m = both & (got_ps == ps)
        same = int((truth_hp[m] == got_hp[m]).sum())
        agree += max(same, int(m.sum()) - same)
        total += int(m.sum())
    return agree, total


def test_phase_library_exports_header_symbols():
    from nanocaller_b200.host import phasing
    lib = phasing.load_phase_library()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "nanocaller_b200_phase.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(nc_[a-z0-9_]+)\s*\(", src)))
    assert names == sorted(phasing.PHASE_EXPORTS)
    for n in names:
        assert hasattr(lib, n)


@pytest.mark.parametrize("preset,cov", [("ont", 30.0), ("hifi", 30.0)])
def test_haplotags_recover_the_true_haplotypes(preset, cov):
    from nanocaller_b200.host import phasing
    w = _world(chrom="chrP", preset=preset, contig_len=400_000, seed=51, coverage=cov, het_every=1000, hom_every=3000)
    rs = w.reads
    truth = rs.hp.copy()
    assert set(np.unique(truth).tolist()) == {1, 2}
    lines = _truth_lines(w)
    out, st = phasing.phase_snp_records(lines, rs, phase_qual_score=10.0)
    assert st["het_sites"] > 250 and st["phased_sites"] >= 0.98 * st["het_sites"]
    assert st["tagged_reads"] >= 0.9 * st["reads"]
    agree, total = _block_agreement(truth, rs.hp, rs.ps)
    assert total >= 0.9 * rs.n and agree >= 0.99 * total, (agree, total)
    assert st["blocks"] <= 5                                   # 12-15 kb reads over hets every ~1 kb chain into (almost) one block
    # records: hets phased with PS appended, homozygous calls untouched, nothing lost
    assert len(out) == len(lines)
    n_ph = 0
    for a, b in zip(lines, out):
        fa, fb = a.rstrip("\n").split("\t"), b.rstrip("\n").split("\t")
        assert fa[:8] == fb[:8]
        if fa[9].startswith("1/1"):
            assert a == b
        elif "|" in fb[9]:
            n_ph += 1
            assert fb[8] == fa[8] + ":PS" and fb[9].split(":")[0] in ("0|1", "1|0") and int(fb[9].rsplit(":", 1)[1]) > 0
            assert fb[9].split(":")[1:-1] == fa[9].split(":")[1:]
    assert n_ph == st["phased_sites"]


def test_phased_genotypes_are_consistent_with_the_reads():
    """Within a block, haplotype 1 of the phased records is the haplotype the HP=1 reads come from: at every phased site most HP=1 reads show
    the allele left of the bar."""
    from nanocaller_b200.host import phasing
    from oracle.indel_oracle import read_token
    w = _world(chrom="chrQ", preset="ont", contig_len=120_000, seed=52, coverage=30.0, het_every=1500, hom_every=0)
    rs = w.reads
    out, st = phasing.phase_snp_records(_truth_lines(w), rs)
    checked = 0
    for ln in out[::6]:
        f = ln.rstrip("\n").split("\t")
        if "|" not in f[9]:
            continue
        p0 = int(f[1]) - 1
        left = f[3] if f[9].startswith("0|") else f[4]
        cov = np.nonzero((rs.pos <= p0) & (rs.ref_end > p0) & (rs.hp == 1))[0]
        toks = [read_token(rs, int(i), p0)[0] for i in cov]
        assert sum(t == left for t in toks) > 0.7 * len(toks), (ln, toks)
        checked += 1
    assert checked >= 10


def test_low_quality_and_isolated_sites_stay_unphased():
    from nanocaller_b200.host import phasing
    w = _world(chrom="chrR", preset="ont", contig_len=100_000, seed=53, coverage=30.0, het_every=1000, hom_every=0)
    rs = w.reads
    lines = _truth_lines(w, qual=5.0)
    out, st = phasing.phase_snp_records(lines, rs, phase_qual_score=10.0)          # indelCaller.py:232: QUAL < cutoff is not phased
    assert out == lines and st["het_sites"] == 0 and int((rs.hp > 0).sum()) == 0
    # one site alone has no partner to be phased against
    out, st = phasing.phase_snp_records(_truth_lines(w)[:1], rs)
    assert st["het_sites"] == 1 and st["phased_sites"] == 0 and out == _truth_lines(w)[:1] and int((rs.hp > 0).sum()) == 0


def test_allele_table_against_the_pileup_restatement():
    """nc_phase_read_alleles vs the oracle's pileup strings on every (read, site) pair of a small contig."""
    from nanocaller_b200.host import phasing
    from oracle.indel_oracle import read_token
    w = _world(chrom="chrT", preset="ont", contig_len=30_000, seed=54, coverage=12.0, het_every=700, hom_every=0, clip_prob=0.5)
    rs = w.reads
    pos1, kinds, alts = w.truth_snps()
    keep = [i for i in range(len(pos1)) if chr(rs.ref[pos1[i] - 1]) in "ACGT"]
    pos0 = (pos1[keep] - 1).astype(np.int32)
    nib = {"A": 1, "C": 2, "G": 4, "T": 8}
    na = np.array([nib[chr(rs.ref[p])] for p in pos0], np.uint8)
    nb = np.array([nib[chr(a)] for a in alts[keep]], np.uint8)
    lib = phasing.load_phase_library()
    first = np.searchsorted(pos0, rs.pos).astype(np.int64)
    last = np.searchsorted(pos0, rs.ref_end).astype(np.int64)
    off = np.zeros(rs.n + 1, np.int64)
    np.cumsum(last - first, out=off[1:])
    al = np.full(int(off[-1]), 7, np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.nc_phase_read_alleles(rs.n, P(rs.pos), P(rs.cigar_off), P(rs.cigar), P(rs.seq_off), P(rs.l_seq), P(rs.seq4), len(pos0),
                                     P(pos0), P(na), P(nb), P(first), P(off), P(al), 3) == 0
    n = 0
    for r in range(rs.n):
        for k in range(int(off[r + 1] - off[r])):
            j = int(first[r]) + k
            t = read_token(rs, r, int(pos0[j]))[0]
            want = 0 if t == chr(rs.ref[pos0[j]]) else (1 if t == chr(alts[keep][j]) else 255)
            assert al[off[r] + k] == want, (r, j, t)
            n += 1
    assert n > 300


def test_cli_phase_stage_from_a_pass_vcf(tmp_path):
    """The command line's phasing stage without the GPU stages around it: PASS SNP file in, `{prefix}.snps.phased.vcf.gz` (+ CSI) out,
    reads of the registered source tagged; contigs whose reads are tagged already are left alone unless --phase is given."""
    import gzip
    from nanocaller_b200 import cli
    from nanocaller_b200.host import sources, vcfio
    w = _world(chrom="chrU", preset="ont", contig_len=150_000, seed=55, coverage=30.0, het_every=1000, hom_every=3000)
    y = _world(chrom="chrY", preset="ont", contig_len=30_000, seed=56, coverage=20.0, het_every=0, hom_every=2000, ploidy=1)
    truth = w.reads.hp.copy()
    sources.unregister_all()
    sources.register_source("mem://p", [w.reads, y.reads])
    lines = _truth_lines(w) + ["chrY\t%d\t.\tA\tC\t60.00\tPASS\t.\tGT:DP:VF:AD:ADF:ADR\t1/1:20:1.0000:0,20:0,10:0,10\n" % p for p in (100, 2000)]
    passp = str(tmp_path / "t.snps.vcf.gz")
    vcfio.write_vcf(passp, "snps", ["chrU", "chrY"], lines, "S", index=True)
    args = cli.parse_args(["--bam", "mem://p", "--ref", "mem://p", "--mode", "all", "--preset", "ont", "--output", str(tmp_path), "--prefix", "t", "--sample", "S"])
    regions = [("chrU", 1, 150_000, "diploid"), ("chrY", 1, 30_000, "haploid")]
    out = {"snps": passp}
    cli._phase_stage(args, regions, ["chrU", "chrY"], out)              # tagged already: nothing to do
    assert out["phase_stats"] == {} and vcfio.read_records(out["phased_snps"]) == vcfio.read_records(passp)
    w.reads.hp[:] = 0
    w.reads.ps[:] = 0
    out = {"snps": passp}
    cli._phase_stage(args, regions, ["chrU", "chrY"], out)
    assert out["phased_snps"].endswith("t.snps.phased.vcf.gz") and os.path.exists(out["phased_snps"] + ".csi")
    st = out["phase_stats"]["chrU"]
    assert st["phased_sites"] >= 0.98 * st["het_sites"] > 100
    with gzip.open(out["phased_snps"], "rt") as f:
        txt = f.read()
    assert '##FORMAT=<ID=PS,Number=1,Type=Integer,Description="Phase set identifier">\n#CHROM' in txt
    recs = [ln for ln in txt.splitlines() if not ln.startswith("#")]
    assert len(recs) == len(lines) and sum("|" in r.split("\t")[9] for r in recs) == st["phased_sites"]
    assert [r for r in recs if r.startswith("chrY")] == [ln.rstrip("\n") for ln in lines if ln.startswith("chrY")]
    agree, total = _block_agreement(truth, w.reads.hp, w.reads.ps)
    assert total >= 0.9 * w.reads.n and agree >= 0.99 * total
    sources.unregister_all()


@pytest.mark.parametrize("preset,cov,het", [("ont", 30.0, 1000), ("ont", 12.0, 1000), ("ont", 30.0, 8000), ("hifi", 15.0, 3000)])
def test_false_heterozygous_calls_do_not_disturb_the_phasing(preset, cov, het):
    """A third more sites that are no variants at all (every read shows the reference base up to sequencing errors): they stay unphased
    and the reads are tagged exactly as without them."""
    from nanocaller_b200.host import phasing
    w = _world(chrom="chrJ", preset=preset, contig_len=1_000_000, seed=61, coverage=cov, het_every=het, hom_every=3000)
    rs = w.reads
    truth = rs.hp.copy()
    lines = _truth_lines(w, hom_too=False)
    _, st0 = phasing.phase_snp_records(lines, rs)
    hp0, ps0 = rs.hp.copy(), rs.ps.copy()
    have = {int(l.split("\t")[1]) for l in lines}
    rng = np.random.RandomState(4)
    junk = []
    for p in rng.choice(np.arange(1000, 999_000), size=max(20, len(lines) // 3), replace=False).tolist():
        ref = chr(rs.ref[p - 1])
        if p in have or ref not in "ACGT":
            continue
        alt = [c for c in "ACGT" if c != ref][rng.randint(3)]
        junk.append("%s\t%d\t.\t%s\t%s\t30.00\tPASS\t.\tGT:DP:VF:AD:ADF:ADR\t0/1:30:0.5000:15,15:8,7:7,8\n" % (rs.chrom, p, ref, alt))
    mixed = sorted(lines + junk, key=lambda l: int(l.split("\t")[1]))
    out, st = phasing.phase_snp_records(mixed, rs)
    assert st["het_sites"] == st0["het_sites"] + len(junk) and abs(st["phased_sites"] - st0["phased_sites"]) <= 0.01 * st0["phased_sites"]
    junk_phased = sum("|" in ln.split("\t")[9] for ln in out if int(ln.split("\t")[1]) not in have)
    assert junk_phased <= 0.02 * len(junk)              # at low coverage two sequencing errors can make a false site look usable
    assert (rs.hp == hp0).mean() >= 0.995
    agree, total = _block_agreement(truth, rs.hp, rs.ps)
    assert total >= (0.8 if het >= 8000 else 0.95) * rs.n and agree >= 0.995 * total, (agree, total, rs.n)


def test_cli_phase_stage_writes_the_haplotagged_bam(tmp_path):
    """--write_phased_bam: intermediate_phase_files/{contig}.phased.bam carries the tags the phasing step gave the reads (indelCaller.py:244)."""
    from nanocaller_b200 import cli
    from nanocaller_b200.host import bamio, sources, vcfio
    w = _world(chrom="chrV", preset="ont", contig_len=80_000, seed=57, coverage=20.0, het_every=1000, hom_every=0)
    rs = w.reads
    lines = _truth_lines(w)
    rs.hp[:] = 0
    rs.ps[:] = 0
    bam, fa = str(tmp_path / "v.bam"), str(tmp_path / "v.fa")
    bamio.write_bam(bam, [rs], index=True)
    bamio.write_fasta(fa, [rs])
    passp = str(tmp_path / "t.snps.vcf.gz")
    vcfio.write_vcf(passp, "snps", ["chrV"], lines, "S", index=True)
    sources.unregister_all()
    bamio.open_alignment(bam, fa, contigs={"chrV"})
    args = cli.parse_args(["--bam", bam, "--ref", fa, "--mode", "all", "--preset", "ont", "--output", str(tmp_path), "--prefix", "t", "--write_phased_bam"])
    out = {"snps": passp}
    cli._phase_stage(args, [("chrV", 1, 80_000, "diploid")], ["chrV"], out)
    pb = out["phase_stats"]["chrV"]["phased_bam"]
    assert pb == str(tmp_path / "intermediate_phase_files" / "chrV.phased.bam")
    mem = sources.resolve(bam, "chrV")
    got = [r for r in bamio.read_bam_native(pb)[0] if r.chrom == "chrV"][0]
    assert int((mem.hp > 0).sum()) > 0.9 * mem.n
    assert np.array_equal(got.hp, mem.hp) and np.array_equal(got.ps, mem.ps) and np.array_equal(got.pos, mem.pos)
    sources.unregister_all()


def test_two_alternative_alleles_are_phased_like_ref_alt():
    """`1/2` calls (both haplotypes non-reference): allele A = ALT1, B = ALT2.  Rewriting 0/1 records (REF r, ALT a) as 1/2 records over a
    third base (REF x, ALT r,a) must give the same phase: `0|1` <-> `1|2`, `1|0` <-> `2|1`, same PS, same read tags."""
    from nanocaller_b200.host import phasing
    w = _world(chrom="chrM2", preset="ont", contig_len=150_000, seed=58, coverage=25.0, het_every=1200, hom_every=0)
    rs = w.reads
    lines = _truth_lines(w)
    out01, st01 = phasing.phase_snp_records(lines, rs)
    hp01 = rs.hp.copy()
    alt_lines = []
    for k, ln in enumerate(lines):
        f = ln.rstrip("\n").split("\t")
        if k % 2 == 0:
            x = [c for c in "ACGT" if c not in (f[3], f[4])][0]
            f[3], f[4] = x, f[3] + "," + f[4]
            f[9] = "1/2" + f[9][3:]
        alt_lines.append("\t".join(f) + "\n")
    out12, st12 = phasing.phase_snp_records(alt_lines, rs)
    assert st12 == st01 and np.array_equal(rs.hp, hp01)
    conv = {"0|1": "1|2", "1|0": "2|1"}
    n = 0
    for k, (a, b) in enumerate(zip(out01, out12)):
        ga, gb = a.rstrip("\n").split("\t")[9], b.rstrip("\n").split("\t")[9]
        if "|" not in ga:
            assert "|" not in gb
            continue
        want = conv[ga[:3]] if k % 2 == 0 else ga[:3]
        assert gb[:3] == want and gb.rsplit(":", 1)[1] == ga.rsplit(":", 1)[1]
        n += 1
    assert n > 80
