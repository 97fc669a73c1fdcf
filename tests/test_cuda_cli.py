"""GPU, end to end through files: synthetic BAM + FASTA on disk -> `python -m nanocaller_b200` (cli.main) -> VCFs, against the
oracle pipeline on the same reads (oracle tensors -> scaling -> fp32 CNN -> restated record code), with the reference's
chunk grid for `--cpu 3`, a diploid and a haploid (chrY) contig, SNP and indel modes."""
import gzip

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _world(tmp_path):
    from nanocaller_b200.host import bamio
    from nanocaller_b200.synth import make_world
    a = make_world(chrom="chrA", preset="ont", contig_len=150_000, seed=11, coverage=24.0, indel_every=1500, indel_maxlen=12).reads
    y = make_world(chrom="chrY", preset="ont", contig_len=40_000, seed=12, coverage=20.0, indel_every=1500, indel_maxlen=12).reads
    bam, fa = str(tmp_path / "w.bam"), str(tmp_path / "w.fa")
    bamio.write_bam(bam, [a, y], index=True)                 # with a BAI the command line reads only the contigs it calls
    bamio.write_fasta(fa, [a, y])
    return bam, fa, [a, y]


def _records(path):
    with gzip.open(path, "rt") as f:
        txt = f.read()
    return [ln + "\n" for ln in txt.splitlines() if not ln.startswith("#")], txt


def test_cli_snps_and_indels_from_files(tmp_path):
    from nanocaller_b200 import cli
    from nanocaller_b200.host import snp_pileups, sources, vcfio, weights as W
    from nanocaller_b200.host.vcf_compare import compare_records
    from oracle import cnn_oracle, indel_caller_oracle, indel_oracle, snp_caller_oracle, snp_oracle
    bam, fa, worlds = _world(tmp_path)
    sources.unregister_all()
    snp_pileups.reset()
    out = cli.main(["--bam", bam, "--ref", fa, "--mode", "all", "--preset", "ont", "--cpu", "3", "--output", str(tmp_path / "o"), "--prefix", "t", "--sample", "S"])
    assert out["launches"] > 0

    # ---- SNPs: oracle pipeline over the reference's chunk grid
    regions = [("chrA", 1, 150_000, "diploid"), ("chrY", 1, 40_000, "haploid")]
    dct = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
    tensors, meta = W.load_model("snp", "ONT-HG002")
    hap, _ = W.load_model("snp", "haploid")
    want = []
    for ch in snp_oracle.get_chunks(regions, 3):
        rs = worlds[0] if ch["chrom"] == "chrA" else worlds[1]
        pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(rs, dct, ch)
        if len(pos) == 0:
            continue
        ref = np.asarray(ref, np.float32)
        if ch["ploidy"] == "haploid":
            x = snp_oracle.scale_counts(mat, 30.0, coverage=float(depth))
            want += snp_caller_oracle.haploid_records(ch["chrom"], pos, ref, cnn_oracle.haploid_snp_model(hap, x, ref), dp, freq)
        else:
            x = snp_oracle.scale_counts(mat, meta["train_coverage"], coverage=float(depth))
            want += snp_caller_oracle.diploid_records(ch["chrom"], pos, ref, cnn_oracle.snp_probs(tensors, x, ref), dp, freq, fwd, rev)
    want = vcfio.sort_records(want, ["chrA", "chrY"])
    got, txt = _records(out["unfiltered_snps"])
    assert txt.startswith("##fileformat=VCFv4.2\n") and "##contig=<ID=chrA>\n##contig=<ID=chrY>\n" in txt and "FORMAT\tS\n" in txt
    res = compare_records(got, want, tol=1e-4)
    assert not res["mismatch"], res["mismatch"][:3]
    assert res["identical"] + res["numeric_only"] + res["borderline"] == len(want) > 500
    assert res["borderline"] <= max(1, len(want) // 500)
    passed, _ = _records(out["snps"])
    assert passed == [ln for ln in got if ln.split("\t")[6] == "PASS"] and 0 < len(passed) < len(got)
    # every output comes with its CSI index (tabix -fp vcf --csi in the reference): a region query through it equals the scan
    import os
    for key in ("unfiltered_snps", "snps", "indels", "final"):
        assert os.path.exists(out[key] + ".csi"), key
    q = vcfio.csi_query(out["unfiltered_snps"], "chrA", 50_000, 60_000)
    assert q == [ln for ln in got if ln.startswith("chrA\t") and 50_000 <= int(ln.split("\t")[1]) - 1 < 60_000] and len(q) > 50

    # ---- indels: oracle pipeline over the 100 kb indel grid (HP / PS tags come from the BAM)
    idct = dict(mincov=4, maxcov=160, seq="ont", del_t=0.6, ins_t=0.4, impute_indel_phase=False, supplementary=False, win_size=40, small_win_size=4)
    it, _ = W.load_model("indel", "ONT-HG002")
    ih, _ = W.load_model("indel", "haploid")
    want_i = []
    for ch in snp_oracle.get_chunks(regions, 3, 100_000):
        rs = worlds[0] if ch["chrom"] == "chrA" else worlds[1]
        if ch["ploidy"] == "haploid":
            pos, x, alleles = indel_oracle.get_indel_testing_candidates_haploid(rs, idct, ch)
            if len(pos):
                want_i += indel_caller_oracle.haploid_records(ch["chrom"], pos, cnn_oracle.haploid_indel_model(ih, np.asarray(x, np.float32)), alleles)
        else:
            pos, x0, x1, x2, alleles, phase = indel_oracle.get_indel_testing_candidates(rs, idct, ch)
            if len(pos):
                probs = cnn_oracle.indel_model(it, np.hstack([x0, x1, x2]).astype(np.float32))
                want_i += indel_caller_oracle.diploid_records(ch["chrom"], pos, probs, alleles, phase)
    want_i = vcfio.sort_records(want_i, ["chrA", "chrY"])
    got_i, txt_i = _records(out["indels"])
    assert "ID=GQ" in txt_i and len(got_i) == len(want_i) > 10
    for a, b in zip(got_i, want_i):
        fa_, fb_ = a.split("\t"), b.split("\t")
        assert fa_[:5] == fb_[:5] and fa_[6:9] == fb_[6:9], (a, b)
        assert abs(float(fa_[5]) - float(fb_[5])) < 0.02
        assert fa_[9].split(":")[0] == fb_[9].split(":")[0]
    merged, txt_m = _records(out["final"])
    assert len(merged) == len(passed) + len(got_i) and "ID=PS" in txt_m and "ID=PR" in txt_m
    keys = [(ln.split("\t")[0], int(ln.split("\t")[1])) for ln in merged]
    assert keys == sorted(keys)


def test_cli_ccs_preset_regions_and_exclude_bed(tmp_path):
    """`--preset ccs --mode snps` (pacbio neighbour rule, CCS-HG002 model, thresholds 0.3,0.7), regions given out of order with an
    explicit sub-range, an exclude BED given by path: records against the oracle pipeline with the same parameters."""
    from nanocaller_b200 import cli
    from nanocaller_b200.host import snp_pileups, sources, vcfio, weights as W
    from nanocaller_b200.host.vcf_compare import compare_records
    from oracle import cnn_oracle, snp_caller_oracle, snp_oracle
    bam, fa, worlds = _world(tmp_path)
    bed = str(tmp_path / "ex.bed")
    with open(bed, "w") as f:
        f.write("chrA\t20000\t26000\nchrA\t90000\t90500\n")
    sources.unregister_all()
    snp_pileups.reset()
    out = cli.main(["--bam", bam, "--ref", fa, "--mode", "snps", "--preset", "ccs", "--cpu", "2", "--regions", "chrY", "chrA:10001-120000",
                    "--exclude_bed", bed, "--output", str(tmp_path / "o2"), "--prefix", "c"])
    regions = [("chrY", 1, 40_000, "haploid"), ("chrA", 10_001, 120_000, "diploid")]
    dct = dict(threshold=[0.3, 0.7], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="pacbio", supplementary=False)
    tensors, meta = W.load_model("snp", "CCS-HG002")
    hap, _ = W.load_model("snp", "haploid")
    ivs = {"chrA": [(20000, 26000), (90000, 90500)]}
    want = []
    for ch in snp_oracle.get_chunks(regions, 2):
        rs = worlds[0] if ch["chrom"] == "chrA" else worlds[1]
        pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(rs, dct, ch, ivs.get(ch["chrom"]))
        if len(pos) == 0:
            continue
        ref = np.asarray(ref, np.float32)
        if ch["ploidy"] == "haploid":
            x = snp_oracle.scale_counts(mat, 30.0, coverage=float(depth))
            want += snp_caller_oracle.haploid_records(ch["chrom"], pos, ref, cnn_oracle.haploid_snp_model(hap, x, ref), dp, freq)
        else:
            x = snp_oracle.scale_counts(mat, meta["train_coverage"], coverage=float(depth))
            want += snp_caller_oracle.diploid_records(ch["chrom"], pos, ref, cnn_oracle.snp_probs(tensors, x, ref), dp, freq, fwd, rev)
    want = vcfio.sort_records(want, ["chrY", "chrA"])                      # header / output order follows the order of --regions
    got, txt = _records(out["unfiltered_snps"])
    assert txt.index("##contig=<ID=chrY>") < txt.index("##contig=<ID=chrA>")
    assert not any(20000 <= int(ln.split("\t")[1]) < 26000 for ln in got if ln.startswith("chrA"))
    res = compare_records(got, want, tol=1e-4)
    assert not res["mismatch"], res["mismatch"][:3]
    assert res["identical"] + res["numeric_only"] + res["borderline"] == len(want) > 300


def test_cli_indels_on_an_untagged_bam(tmp_path):
    """No HP tags: both haplotype depths stay below mincov (generate_indel_pileups.py:252).  Without impute_indel_phase the reference finds
    no candidates and the command line must come back with header-only, indexed outputs; with it (`--preset ccs` sets the flag,
    NanoCaller:74) the read sets come from the pileup strings (:278-304) and the records must equal the oracle pipeline's."""
    import os
    from nanocaller_b200 import cli
    from nanocaller_b200.host import bamio, snp_pileups, sources, vcfio, weights as W
    from nanocaller_b200.synth import make_world
    from oracle import cnn_oracle, indel_caller_oracle, indel_oracle, snp_oracle
    rs = make_world(chrom="chrN", preset="hifi", contig_len=60_000, seed=13, coverage=30.0, indel_every=1500, indel_maxlen=12, untagged_frac=1.0).reads
    assert int((rs.hp > 0).sum()) == 0
    bam, fa = str(tmp_path / "u.bam"), str(tmp_path / "u.fa")
    bamio.write_bam(bam, [rs], index=True)
    bamio.write_fasta(fa, [rs])
    sources.unregister_all()
    snp_pileups.reset()
    out = cli.main(["--bam", bam, "--ref", fa, "--mode", "indels", "--preset", "ont", "--output", str(tmp_path / "o3")])
    assert out["n_indel_records"] == 0
    recs, txt = _records(out["indels"])
    assert recs == [] and txt.startswith("##fileformat=VCFv4.2\n") and "##contig=<ID=chrN>" in txt
    assert os.path.exists(out["indels"] + ".csi")

    sources.unregister_all()
    snp_pileups.reset()
    out = cli.main(["--bam", bam, "--ref", fa, "--mode", "indels", "--preset", "ccs", "--cpu", "2", "--output", str(tmp_path / "o4")])
    idct = dict(mincov=4, maxcov=160, seq="pacbio", del_t=0.4, ins_t=0.4, impute_indel_phase=True, supplementary=False, win_size=40, small_win_size=4)
    it, _ = W.load_model("indel", "CCS-HG002")
    want = []
    for ch in snp_oracle.get_chunks([("chrN", 1, 60_000, "diploid")], 2, 100_000):
        pos, x0, x1, x2, alleles, phase = indel_oracle.get_indel_testing_candidates(rs, idct, ch)
        if len(pos):
            probs = cnn_oracle.indel_model(it, np.hstack([x0, x1, x2]).astype(np.float32))
            want += indel_caller_oracle.diploid_records(ch["chrom"], pos, probs, alleles, phase)
    want = vcfio.sort_records(want, ["chrN"])
    got, _ = _records(out["indels"])
    assert len(got) == len(want) > 10
    for a, b in zip(got, want):
        fa_, fb_ = a.split("\t"), b.split("\t")
        assert fa_[:5] == fb_[:5] and fa_[6:9] == fb_[6:9], (a, b)
        assert abs(float(fa_[5]) - float(fb_[5])) < 0.02
        assert fa_[9].split(":")[0] == fb_[9].split(":")[0]


def test_cli_mode_all_phases_an_untagged_bam(tmp_path):
    """`--mode all` on a BAM without HP tags: the SNP calls are phased and the reads haplotagged in memory (host/phasing.py, in place of
    `whatshap phase` / `haplotag`, indelCaller.py:237,:244), so the indel stage finds (nearly) the same variants as on the same reads
    carrying the generator's true haplotype tags; the phased SNP file and the merged file carry the phased genotypes."""
    import copy
    import os
    from nanocaller_b200 import cli
    from nanocaller_b200.host import bamio, snp_pileups, sources
    from nanocaller_b200.synth import make_world
    rs = make_world(chrom="chrB", preset="ont", contig_len=300_000, seed=21, coverage=30.0, indel_every=1500, indel_maxlen=12).reads
    truth_hp = rs.hp.copy()
    assert set(np.unique(truth_hp).tolist()) == {1, 2}
    un = copy.copy(rs)
    un.hp, un.ps = np.zeros_like(rs.hp), np.zeros_like(rs.ps)
    fa = str(tmp_path / "b.fa")
    bamio.write_fasta(fa, [rs])
    runs = {}
    for tag, r in (("tagged", rs), ("untagged", un)):
        bam = str(tmp_path / (tag + ".bam"))
        bamio.write_bam(bam, [r], index=True)
        sources.unregister_all()
        snp_pileups.reset()
        out = cli.main(["--bam", bam, "--ref", fa, "--mode", "all", "--preset", "ont", "--cpu", "2", "--output", str(tmp_path / tag)])
        runs[tag] = (out, _records(out["indels"])[0], _records(out["phased_snps"])[0], sources.resolve(bam, "chrB"))
    out_t, ind_t, ph_t, _ = runs["tagged"]
    out_u, ind_u, ph_u, rs_u = runs["untagged"]
    # tagged BAM: the tags are used as they are, nothing is phased
    assert out_t["phase_stats"] == {} and not any("|" in ln.split("\t")[9] for ln in ph_t)
    # untagged BAM: phased hets with PS, reads tagged like the truth (up to the naming of the two haplotypes per block)
    st = out_u["phase_stats"]["chrB"]
    assert st["het_sites"] > 150 and st["phased_sites"] >= 0.95 * st["het_sites"] and st["tagged_reads"] >= 0.85 * st["reads"]
    n_ph = sum("|" in ln.split("\t")[9] for ln in ph_u)
    assert n_ph == st["phased_sites"] and all(ln.split("\t")[8].endswith(":PS") for ln in ph_u if "|" in ln.split("\t")[9])
    assert len(ph_u) == len(ph_t) and [ln.split("\t")[:5] for ln in ph_u] == [ln.split("\t")[:5] for ln in ph_t]
    both = (truth_hp > 0) & (rs_u.hp > 0)
    agree = total = 0
    for ps in np.unique(rs_u.ps[both]):
        m = both & (rs_u.ps == ps)
        same = int((truth_hp[m] == rs_u.hp[m]).sum())
        agree += max(same, int(m.sum()) - same); total += int(m.sum())
    assert total >= 0.85 * rs.n and agree >= 0.97 * total, (agree, total)
    # indel calls: the same variants as with the true tags
    key = lambda ln: tuple(ln.split("\t")[:2]) + (ln.split("\t")[3],) + (tuple(sorted(ln.split("\t")[4].split(","))),)
    kt, ku = {key(ln) for ln in ind_t}, {key(ln) for ln in ind_u}
    assert len(kt) > 100 and len(kt & ku) >= 0.9 * len(kt) and len(ku) <= 1.1 * len(kt), (len(kt), len(ku), len(kt & ku))
    merged = _records(out_u["final"])[0]
    assert len(merged) == len(ph_u) + len(ind_u) and os.path.exists(out_u["phased_snps"] + ".csi")


def test_cli_decompose_indels_flag(tmp_path):
    """`--decompose_indels`: the records of the indel stage go to intermediate_indel_files/{prefix}.raw.indel.vcf (indelCaller.py:369) and
    `{prefix}.indels.vcf.gz` holds their normalised form (host/vcf_decompose.py, in place of indelCaller.py:391)."""
    from nanocaller_b200 import cli
    from nanocaller_b200.host import snp_pileups, sources, vcfio
    from nanocaller_b200.host.vcf_decompose import decompose_records
    bam, fa, worlds = _world(tmp_path)
    outs = {}
    for tag, extra in (("plain", []), ("dec", ["--decompose_indels"])):
        sources.unregister_all()
        snp_pileups.reset()
        outs[tag] = cli.main(["--bam", bam, "--ref", fa, "--mode", "indels", "--preset", "ont", "--cpu", "3", "--regions", "chrA",
                              "--output", str(tmp_path / tag)] + extra)
    plain = vcfio.read_records(outs["plain"]["indels"])
    raw = vcfio.read_records(outs["dec"]["raw_indels"])
    dec = vcfio.read_records(outs["dec"]["indels"])
    assert raw == plain and len(plain) > 20 and "raw_indels" not in outs["plain"]
    assert dec == decompose_records(plain, contigs=["chrA"]) and dec != plain
    assert sum(len(ln.split("\t")[3]) for ln in dec) < sum(len(ln.split("\t")[3]) for ln in plain)      # padding context is gone
