"""CPU: the command-line surface of the reference (NanoCaller:84-174, utils.py:6-83) as mirrored by nanocaller_b200/cli.py:
flags and defaults, the preset rule, region / ploidy rules and the chunk grid (against the oracle's pinned get_chunks)."""
import numpy as np
import pytest

from nanocaller_b200 import cli
from oracle import snp_oracle


def test_defaults_match_reference_parser():
    a = cli.parse_args(["--bam", "x.bam", "--ref", "x.fa"])
    assert (a.mode, a.sequencing, a.cpu, a.mincov, a.maxcov) == ("all", "ont", 1, 4, 160)
    assert (a.snp_model, a.indel_model, a.min_allele_freq, a.min_nbr_sites, a.neighbor_threshold) == ("ONT-HG002", "ONT-HG002", 0.15, 1, "0.4,0.6")
    assert (a.ins_threshold, a.del_threshold, a.win_size, a.small_win_size) == (0.4, 0.6, 40, 4)
    assert (a.prefix, a.sample, a.phase_qual_score, a.supplementary) == ("variant_calls", "SAMPLE", 10, False)


def test_preset_applies_unless_flag_given():
    a = cli.parse_args(["--bam", "x", "--ref", "y", "--preset", "ccs"])
    assert (a.sequencing, a.snp_model, a.indel_model, a.neighbor_threshold, a.ins_threshold, a.del_threshold) == ("pacbio", "CCS-HG002", "CCS-HG002", "0.3,0.7", 0.4, 0.4)
    assert a.impute_indel_phase and a.enable_whatshap
    a = cli.parse_args(["--bam", "x", "--ref", "y", "--preset", "clr", "--win_size", "25", "--snp_model", "ONT-HG001"])
    assert (a.win_size, a.small_win_size, a.snp_model, a.indel_model) == (25, 2, "ONT-HG001", "ONT-HG002")
    with pytest.raises(SystemExit):
        cli.parse_args(["--bam", "x"])                       # --ref is required


def test_regions_and_ploidy_rules():
    lens = {"chr1": 1000, "chrX": 500, "chrY": 300, "chrM": 16, "weird": 50}
    a = cli.parse_args(["--bam", "x", "--ref", "y"])
    r = cli.get_regions_list(a, lens)
    assert r == [("chr1", 1, 1000, "diploid"), ("chrX", 1, 500, "diploid"), ("chrY", 1, 300, "haploid"), ("chrM", 1, 16, "haploid"), ("weird", 1, 50, "diploid")]
    a = cli.parse_args(["--bam", "x", "--ref", "y", "--haploid_X", "--regions", "chrX", "chr1:10-200", "nope", "chr1:5"])
    assert cli.get_regions_list(a, lens) == [("chrX", 1, 500, "haploid"), ("chr1", 10, 200, "diploid")]
    a = cli.parse_args(["--bam", "x", "--ref", "y", "--wgs_contigs", "chr1-22XY", "--haploid_genome"])
    assert cli.get_regions_list(a, lens) == [("chr1", 1, 1000, "haploid"), ("chrX", 1, 500, "diploid"), ("chrY", 1, 300, "haploid")]
    a = cli.parse_args(["--bam", "x", "--ref", "y", "--regions", "absent"])
    with pytest.raises(SystemExit):
        cli.get_regions_list(a, lens)


def test_chunk_grid_equals_pinned_oracle():
    rng = np.random.RandomState(0)
    for _ in range(50):
        regions = []
        for i in range(rng.randint(1, 4)):
            s = int(rng.randint(1, 50_000))
            regions.append(("c%d" % i, s, s + int(rng.randint(1, 3_000_000)), "diploid"))
        cpu = int(rng.randint(1, 40))
        for mx in (500_000, 100_000):
            assert cli.get_chunks(regions, cpu, mx) == snp_oracle.get_chunks(regions, cpu, mx)


def test_product_cli_does_not_import_the_oracle():
    import ast
    import inspect
    tree = ast.parse(inspect.getsource(cli))
    names = [n.module or "" for n in ast.walk(tree) if isinstance(n, ast.ImportFrom)] + [a.name for n in ast.walk(tree) if isinstance(n, ast.Import) for a in n.names]
    assert not any(n.startswith("oracle") for n in names)


def test_every_reference_flag_is_accepted_with_the_reference_default():
    """tests/golden/reference_cli_flags.txt = the add_argument names and defaults of the reference script (NanoCaller:84-158, extracted
    by tests/golden/make_cli_flags.py): the parser knows every flag, has the same default for it, and its own additions are the
    documented four."""
    import ast
    import os
    from nanocaller_b200 import cli
    rows = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cli_flags.txt"))]
    want = {k: ast.literal_eval(v) for k, v in rows}
    acts = {s: a for a in cli.build_parser()._actions for s in a.option_strings if s.startswith("--")}
    assert len(want) > 30 and set(want) <= set(acts), sorted(set(want) - set(acts))
    for k, v in want.items():
        assert acts[k].default == v, (k, acts[k].default, v)
    assert set(acts) - set(want) - {"--help"} == {"--device", "--nanocaller_src", "--write_phased_bam", "--decompose_indels", "--host_bam_reader"}


def test_preset_table_equals_the_reference():
    """tests/golden/reference_presets.json = `preset_dict` of the reference script (NanoCaller:66-77, extracted by make_cli_flags.py)."""
    import json
    import os
    from nanocaller_b200 import cli
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_presets.json")))
    assert cli.PRESETS == want and len(want) == 6


def test_model_name_tables_equal_the_reference():
    """tests/golden/reference_model_tables.json = snp_model_dict / indel_model_dict of the reference (snpCaller.py:16-34, indelCaller.py:17-24)."""
    import json
    import os
    from nanocaller_b200.host import weights
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model_tables.json")))
    assert weights.SNP_MODEL_DICT == want["snp"] and weights.INDEL_MODEL_DICT == want["indel"]


def test_regions_list_equals_the_unmodified_reference(tmp_path, capsys):
    """tests/golden/reference_regions.json = answers (regions, exit codes, messages) of the unmodified `utils.get_regions_list`
    (utils.py:6-65) run over the pysam shim (tests/golden/make_regions_golden.py) on eleven argument scenarios."""
    import argparse
    import json
    import os
    import pytest
    from nanocaller_b200 import cli
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_regions.json")))
    bed = tmp_path / "r.bed"
    bed.write_text(g["bed"])
    assert len(g["scenarios"]) >= 11
    for sc in g["scenarios"]:
        lengths = dict((n, l) for n, l in g["contigs"][sc["source"]])
        args = argparse.Namespace(wgs_contigs=sc.get("wgs_contigs"), regions=sc.get("regions"), bed=str(bed) if sc.get("bed") else None,
                                  haploid_genome=sc.get("haploid_genome", False), haploid_X=sc.get("haploid_X", False))
        capsys.readouterr()
        if isinstance(sc["result"], dict):
            with pytest.raises(SystemExit) as e:
                cli.get_regions_list(args, lengths)
            assert e.value.code == sc["result"]["exit"], sc["name"]
        else:
            assert [list(r) for r in cli.get_regions_list(args, lengths)] == sc["result"], sc["name"]
        msgs = [ln.split(": ", 1)[1] for ln in capsys.readouterr().out.splitlines() if ": " in ln]
        assert msgs == sc["messages"], sc["name"]
