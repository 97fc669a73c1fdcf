"""Indel golden-case definitions (see cases.py).  Reads carry synthetic HP/PS tags (whatshap bypass)."""
from nanocaller_b200.synth import make_world

BASE_INDEL_DCT = dict(mincov=4, maxcov=160, seq="ont", supplementary=False, exclude_bed=None, win_size=40, small_win_size=4,
                      ins_t=0.4, del_t=0.6, impute_indel_phase=False)


def _world(**kw):
    return make_world(**kw).reads


INDEL_CASES = {
    # name: (readset factory, dct overrides, regions, cpu, max_chunk_size)
    "indel_ont": (lambda: _world(chrom="chr20", preset="ont", contig_len=60_000, seed=40, coverage=30.0, indel_every=1500, indel_maxlen=12,
                                 het_every=0, hom_every=0, sys_per_10k=0),
                  {}, [("chr20", 1, 60_000, "diploid")], 2, 100000),
    "indel_hifi": (lambda: _world(chrom="chr1", preset="hifi", contig_len=40_000, seed=41, coverage=32.0, indel_every=2000, indel_maxlen=30,
                                  het_every=0, hom_every=0, sys_per_10k=0, untagged_frac=0.1),
                   {"seq": "pacbio", "ins_t": 0.4, "del_t": 0.4}, [("chr1", 1, 40_000, "diploid")], 1, 100000),
    "indel_sub": (lambda: _world(chrom="chr2", preset="ont", contig_len=50_000, seed=42, coverage=24.0, indel_every=900, indel_maxlen=45,
                                 het_every=700, hom_every=0, sys_per_10k=20, clip_prob=0.3),
                  {"win_size": 20, "small_win_size": 2}, [("chr2", 10_001, 40_000, "diploid")], 2, 100000),
    "indel_haploid": (lambda: _world(chrom="chrX", preset="ont", contig_len=40_000, seed=43, coverage=30.0, indel_every=1200, indel_maxlen=20,
                                     het_every=0, hom_every=0, sys_per_10k=0, ploidy=1),
                      {"del_t": 0.5}, [("chrX", 1, 40_000, "haploid")], 1, 100000),
    # impute_indel_phase (generate_indel_pileups.py:278-304): most reads carry no HP tag, so most columns lack phased coverage and
    # the read sets come from grouping the pileup strings of the column
    "indel_impute_hifi": (lambda: _world(chrom="chr3", preset="hifi", contig_len=40_000, seed=44, coverage=32.0, indel_every=800, indel_maxlen=30,
                                         het_every=0, hom_every=0, sys_per_10k=0, untagged_frac=0.8),
                          {"seq": "pacbio", "ins_t": 0.4, "del_t": 0.4, "impute_indel_phase": True}, [("chr3", 1, 40_000, "diploid")], 2, 100000),
    "indel_impute_ont": (lambda: _world(chrom="chr4", preset="ont", contig_len=30_000, seed=45, coverage=30.0, indel_every=700, indel_maxlen=12,
                                        het_every=0, hom_every=0, sys_per_10k=0, untagged_frac=1.0),
                         {"ins_t": 0.3, "del_t": 0.4, "impute_indel_phase": True}, [("chr4", 1, 30_000, "diploid")], 1, 100000),
}


def indel_case_inputs(name):
    factory, over, regions, cpu, mcs = INDEL_CASES[name]
    dct = dict(BASE_INDEL_DCT)
    dct.update(over)
    return factory(), dct, regions, cpu, mcs
