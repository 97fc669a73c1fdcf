"""CPU: indel record normalisation (host/vcf_decompose.py, in place of `rtg vcfdecompose | rtg vcffilter --non-snps-only`,
indelCaller.py:391).  The rule set is this package's own; the tests pin its stated properties."""
import numpy as np

from nanocaller_b200.host.vcf_decompose import components, decompose_records


def _apply(pos, ref, comps):
    """Apply components (1-based positions relative to the record) to REF -> ALT."""
    out, cur = [], pos
    for p, r, a, _ in sorted(comps):
        assert p >= cur, (p, cur, comps)                  # components do not overlap
        out.append(ref[cur - pos:p - pos])
        assert ref[p - pos:p - pos + len(r)] == r, (p, r, ref)
        out.append(a)
        cur = p + len(r)
    out.append(ref[cur - pos:])
    return "".join(out)


def test_components_rebuild_the_alt_allele():
    rng = np.random.RandomState(9)
    n_complex = n_whole = n_delins = 0
    for _ in range(400):
        ref = "".join(rng.choice(list("AGTC"), rng.randint(1, 40)))
        alt = list(ref)
        for _ in range(rng.randint(1, 4)):
            k = rng.randint(0, max(1, len(alt)))
            r = rng.rand()
            if r < 0.4 and len(alt) > 1:
                del alt[k:k + rng.randint(1, 6)]
            elif r < 0.8:
                alt[k:k] = list(rng.choice(list("AGTC"), rng.randint(1, 6)))
            elif alt:
                alt[min(k, len(alt) - 1)] = rng.choice(list("AGTC"))
        alt = "".join(alt) or "A"
        comps = components(1000, ref, alt)
        if ref == alt:
            assert comps == []
            continue
        assert _apply(1000, ref, comps) == alt, (ref, alt, comps)
        for p, r, a, kind in comps:
            assert r and a and r != a
            if kind == "complex":                                                   # a difference opening with a gap at the first base: kept whole
                assert len(comps) == 1 and p == 1000 and r[0] != a[0] and r[-1] != a[-1]
                n_whole += 1
            elif kind == "indel":
                assert r[0] == a[0] or r[-1] == a[-1]                              # anchored
                n_delins += min(len(r), len(a)) > 1                                 # adjacent deletion + insertion stay one component
            else:
                assert len(r) == len(a) and all(x != y for x, y in zip(r, a)) and (kind == "snp") == (len(r) == 1)
        n_complex += len(comps) > 1
    assert n_complex > 50 and n_whole < 40 and n_delins < 60


def test_padded_alleles_of_the_indel_stage_become_minimal():
    assert components(4469, "TACGGGGCTTCCTCA", "TACGGGGCTTC") == [(4479, "CCTCA", "C", "indel")]
    assert components(5922, "AGTAGGAGCGG", "AGTAGGAGCGGTCACTGAAGA") == [(5932, "G", "GTCACTGAAGA", "indel")]
    assert components(100, "CAAAT", "CAAT") == [(100, "CA", "C", "indel")]            # suffix first: homopolymer indels end up left-aligned
    assert components(100, "ACGT", "ACTT") == [(102, "G", "T", "snp")]
    assert components(100, "ACGT", "ATTT") == [(101, "CG", "TT", "mnp")]


def test_records_genotypes_and_the_non_snp_filter():
    L = ["chr1\t100\t.\tTACGGGGCTTCCTCA\tTACGGGGCTTC,TACGGGGCTTCCTCAGG\t20.00\tPASS\t.\tGT:GQ:PS\t1|2:12.00:55\n",
         "chr1\t90\t.\tAC\tA\t21.00\tPASS\t.\tGT:GQ\t1/1:13.00\n",
         "chr1\t300\t.\tCA\tCAA,CAAA\t22.00\tPASS\t.\tGT:GQ\t1|2:14.00\n",
         "chr1\t400\t.\tACGTAC\tACTTAC\t23.00\tPASS\t.\tGT:GQ\t0|1:15.00\n",           # a SNP in indel clothing: dropped
         "chr1\t500\t.\tGTTTA\tGTTA,GTTA\t24.00\tPASS\t.\tGT:GQ\t2|1:16.00\n",          # the same component on both alleles
         "chr0\t50\t.\tAT\tA\t25.00\tPASS\t.\tGT:GQ:PS\t1|0:17.00:7\n"]
    got = decompose_records(L, contigs=["chr0", "chr1"])
    assert got == ["chr0\t50\t.\tAT\tA\t25.00\tPASS\t.\tGT:GQ:PS\t1|0:17.00:7\n",
                   "chr1\t90\t.\tAC\tA\t21.00\tPASS\t.\tGT:GQ\t1/1:13.00\n",
                   "chr1\t110\t.\tCCTCA\tC\t20.00\tPASS\t.\tGT:GQ:PS\t1|0:12.00:55\n",
                   "chr1\t114\t.\tA\tAGG\t20.00\tPASS\t.\tGT:GQ:PS\t0|1:12.00:55\n",
                   "chr1\t300\t.\tC\tCA,CAA\t22.00\tPASS\t.\tGT:GQ\t1|2:14.00\n",
                   "chr1\t500\t.\tGT\tG\t24.00\tPASS\t.\tGT:GQ\t1|1:16.00\n"]
    assert len(decompose_records(L, keep_snps=True)) == len(got) + 1
    # records that are minimal already pass through unchanged, and the operation is idempotent
    assert decompose_records(got, contigs=["chr0", "chr1"]) == got


def test_decomposing_the_indel_stage_records_of_a_synthetic_contig_keeps_every_call():
    """Records as the indel stage writes them (oracle pipeline on a golden case): every input record yields at least one output
    record unless it hides a SNP only, positions stay inside the input record's span, output is sorted."""
    import json
    import os
    from oracle import indel_caller_oracle
    from tests.test_indel_oracle_golden import load_indel_case
    rs, dct, chunks, g = load_indel_case("indel_ont")
    lines = []
    rng = np.random.RandomState(2)
    for ci in range(len(chunks)):
        pos = g["c%d_pos" % ci]
        alleles = json.loads(str(g["c%d_alleles" % ci]))
        phase = json.loads(str(g["c%d_phase" % ci]))
        probs = rng.dirichlet([0.3, 1, 1, 1], len(pos)).astype(np.float32)
        lines += indel_caller_oracle.diploid_records("chr20", pos, probs, alleles, phase)
    assert len(lines) > 20
    got = decompose_records(lines, contigs=["chr20"])
    spans = [(int(f[1]), int(f[1]) + len(f[3])) for f in (ln.split("\t") for ln in lines)]
    for ln in got:
        f = ln.split("\t")
        p = int(f[1])
        assert any(a <= p < b for a, b in spans) and len(f) == 10 and f[3] != f[4]
    assert [int(ln.split("\t")[1]) for ln in got] == sorted(int(ln.split("\t")[1]) for ln in got)
    assert len(got) >= 0.9 * len(lines)


def test_cli_flag_is_off_by_default():
    from nanocaller_b200 import cli
    a = cli.parse_args(["--bam", "x.bam", "--ref", "x.fa", "--preset", "ont"])
    assert a.decompose_indels is False
    assert cli.parse_args(["--bam", "x.bam", "--ref", "x.fa", "--preset", "ont", "--decompose_indels"]).decompose_indels is True
