"""N3: VCF <-> haplotype + CIGAR around realign_hap (vcf.py:36-426 of the reference), without pysam."""
import numpy as np
import pytest

import oracle
from npore_b200 import cig, vcfio

# the five records / two contigs of the reference's test/test_std_vcf.vcf and test/test_std_ref.fasta
_VCF = """##fileformat=VCFv4.2
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE
chr18\t1\t.\tA\tG\t60\tPASS\t.\tGT\t1|1
chr18\t3\t.\tA\tACCCTA\t60\tPASS\t.\tGT\t0|1
chr19\t1\t.\tC\tCAA\t60\tPASS\t.\tGT\t1|1
chr19\t5\t.\tG\tGAG\t60\tPASS\t.\tGT\t1|0
chr19\t15\t.\tC\tG\t60\tPASS\t.\tGT\t1|1
"""
_REFS = {"chr18": "ACAGCGCTATCAGCAGCTAGCATCAGCATCAG", "chr19": "CAAAGAGGGATTTTCAGCCGCGCAACGAGCAG"}


def test_split_and_apply_match_reference(golden, tmp_path):
    """split_vcf + apply_vcf give the haplotype sequences / CIGARs the reference's own functions gave (golden KATs)."""
    p = tmp_path / "t.vcf"
    p.write_text(_VCF)
    recs = vcfio.read_vcf(str(p))
    assert len(recs) == 5 and recs[1].gt == (0, 1) and recs[1].alts == ("ACCCTA",)
    h1, h2 = vcfio.split_vcf(recs)
    data = {(d[0], d[1]): d for d in vcfio.apply_vcf(h1, 1, _REFS) + vcfio.apply_vcf(h2, 2, _REFS)}
    for k in golden("std_vcf_kats.json"):
        d = data[(k["contig"], k["hap"])]
        assert d[2] == k["seq"] and d[3] == k["ref"] and cig.collapse_cigar(d[4]) == k["cigar"]


def test_gen_vcf_and_merge():
    ref = "ACGTACGTAC"
    seq, cg = "ACGACGTAGC", "===" + "D" + "====" + "=" + "I" + "="      # T@3 deleted, G inserted after A@8
    recs = vcfio.gen_vcf([("c", 1, seq, ref, cg)])
    assert recs == [vcfio.Record("c", 3, "GT", ("G",), 60.0, None), vcfio.Record("c", 9, "A", ("AG",), 60.0, None)]
    sub = vcfio.gen_vcf([("c", 2, "AGGT", "ACGT", "=X=="), ("c", 2, "AGGT", "ACGT", "MMMM")])
    assert sub == [vcfio.Record("c", 2, "C", ("G",), 60.0, None)] * 2
    m = vcfio.merge_vcfs(recs, [recs[0], vcfio.Record("c", 9, "A", ("AT",), 60.0, None)])
    assert [(r.pos, r.gt) for r in m] == [(3, (1, 1)), (9, (1, 0)), (9, (0, 1))]
    # round trip: applying the generated records reproduces the haplotype
    assert vcfio.apply_vcf(recs, 1, {"c": ref})[0][2] == seq


def test_apply_overlapping_deletion_rules():
    """vcf.py:226-238: a variant that starts inside the previous deletion."""
    ref = {"c": "AAACCCGGGTTT"}
    recs = [vcfio.Record("c", 3, "ACCC", ("A",), 60.0, None),            # deletes CCC (ref_ptr -> 6)
            vcfio.Record("c", 5, "C", ("CTT",), 60.0, None),             # insertion inside the deletion: allowed
            vcfio.Record("c", 6, "CG", ("C",), 60.0, None)]              # deletion whose first base overlaps: extends it
    (d,) = vcfio.apply_vcf(recs, 1, ref)
    assert cig.collapse_cigar(d[4]) == "3=3D2I1D5=" and d[2] == "AAATTGGTTT"


@pytest.mark.gpu
def test_standardize_vcf_end_to_end(tables, golden, tmp_path):
    """standardize_vcf.py:10-43 on the reference's tiny VCF: the merged records are exactly those derived from the
    reference's own realign_hap outputs (golden KATs)."""
    from npore_b200 import bam, cfg
    cfg.args.sub_scores, cfg.args.np_scores = tables
    cfg.args.max_n, cfg.args.max_l = 6, 100
    p = tmp_path / "t.vcf"
    p.write_text(_VCF)
    merged = vcfio.standardize_vcf(str(p), _REFS, str(tmp_path / "std.vcf"))
    assert vcfio.read_vcf(str(tmp_path / "std.vcf")) == merged                 # text round trip of the writer
    # (an insertion at contig position 0 is dropped by gen_vcf, vcf.py:363 `if ref_ptr > 0 and seq_ptr > 0`, as in the
    #  reference, so apply(gen(x)) == x only holds away from the contig start)
    # the records are those gen_vcf derives from the reference's standardised CIGARs
    want1 = vcfio.gen_vcf([(k["contig"], 1, k["seq"], k["ref"], k["out"]) for k in golden("std_vcf_kats.json") if k["hap"] == 1])
    want2 = vcfio.gen_vcf([(k["contig"], 2, k["seq"], k["ref"], k["out"]) for k in golden("std_vcf_kats.json") if k["hap"] == 2])
    assert merged == vcfio.merge_vcfs(want1, want2, list(_REFS))
