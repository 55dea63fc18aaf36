"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref).

Run where /root/reference exists:   python tests/golden/make_golden.py
Outputs (all in tests/golden/):
  tables.npz         sub_scores[5,5], np_scores[6,101,101] from the reference's calc_score_matrices
                     on guppy5_stats (aln.pyx:62-96; realign.py:87-93).  Opaque inputs of the path.
  golden_sam.json    the reference's one known-answer fixture: read tuples from test/data/reads.sam +
                     ref.fasta (as bam.pyx:34-47 would yield them) and the expected records of
                     test/data/npore_realigned.sam; plus what the compiled reference prints today.
  align_kats.json    test/align.py:20-39 cases -> reference align() output + chunk scores, at
                     (max_b_rows=20, r=10) [test/align.py:59-60] and at defaults.
  np_info_kats.json  test/get_np_info.py:13-18 sequences + the aln.pyx:182-194 docstring example.
  std_vcf_kats.json  realign_hap on the haplotypes of test/test_std_vcf.vcf x test_std_ref.fasta.
  fuzz.json.gz       seeded differential-fuzz cases (npore_b200/synth.py:fuzz_case) with the reference's
                     expanded CIGAR, per-chunk fp32 scores and realign_read CIGAR column.
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402
from npore_b200 import synth  # noqa: E402

REFROOT = "/root/reference"


def ref_standardized(ref, aligned, ir, iq):
    """bam.pyx:65-78 driven with the reference's own functions."""
    c = aligned.replace("X", "M").replace("=", "M")
    if not c:
        return ""
    ic = ref.cig.cig_to_int(c)
    b1, b2 = np.zeros(len(c), np.uint8), np.zeros(len(c), np.uint8)
    ic = ref.cig.push_indels_left(ic, ir, b1, b2, 2)
    ic = ref.cig.push_inss_thru_dels(ic)
    ic = ref.cig.push_indels_left(ic, iq, b1, b2, 1)
    ic = ref.cig.push_inss_thru_dels(ic)
    return ref.cig.int_to_cig(np.asarray(ic)).replace("ID", "M")


def main():
    ref = ref_loader.load_reference(out_prefix="/tmp/npore_golden")
    S, NP = ref_loader.reference_tables(ref)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), sub_scores=S, np_scores=NP)

    # ---- golden SAM
    fasta = "".join(l.strip() for l in open(f"{REFROOT}/test/data/ref.fasta") if not l.startswith(">")).upper()
    reads, expected = [], []
    for line in open(f"{REFROOT}/test/data/reads.sam"):
        if line.startswith("@"):
            continue
        f = line.rstrip("\n").split("\t")
        name, flag, rname, pos, mapq, cigar, seq, qual = f[0], int(f[1]), f[2], int(f[3]), int(f[4]), f[5], f[9], f[10]
        hap = 0
        for tag in f[11:]:
            if tag.startswith("HP:i:"):
                hap = int(tag[5:])
        start = pos - 1
        stop = start + ref.cig.ref_len(ref.cig.expand_cigar(cigar))
        # query_alignment_sequence: soft clips stripped (none in this fixture, asserted)
        assert "S" not in cigar and "H" not in cigar
        reads.append([name, flag, rname, start, mapq, cigar, stop, seq.upper(), qual, fasta[start:stop], hap])
    for line in open(f"{REFROOT}/test/data/npore_realigned.sam"):
        if not line.startswith("@"):
            expected.append(line.rstrip("\n"))
    if os.path.exists("/tmp/npore_golden.sam"):
        os.remove("/tmp/npore_golden.sam")
    for rd in reads:
        ref.bam.realign_read(tuple(rd))
    produced = [l.rstrip("\n") for l in open("/tmp/npore_golden.sam")]
    exp_by_name = {l.split("\t")[0]: l for l in expected}
    assert all(exp_by_name[p.split("\t")[0]] == p for p in produced), "reference no longer reproduces its golden SAM"
    json.dump({"reads": reads, "expected_sam": [exp_by_name[r[0]] for r in reads]},
              open(os.path.join(HERE, "golden_sam.json"), "w"), indent=0)

    # ---- align KATs (test/align.py:20-39)
    cases = [("ACCAGGCAT", "ACCAGGCAT", "9="), ("ACCAGGCAT", "ACAGGCA", "2=1D5=1D"), ("ACCAGGCAT", "ACCCAGGAT", "1=1I5=1D2="),
             ("AAAACCAGGCA", "AAACCAGGCA", "1D10="), ("TAAACCAGGCA", "AAACCAGGCA", "1D10="), ("AAAACCAGGCA", "AAAAACCAGGCA", "1I11="),
             ("AAAACCAGGCA", "TAAAACCAGGCA", "1I11="), ("CCAAAAAATTTTTCC", "CCAAAAATTTTTTCC", "7=1X7="),
             ("CACACACATATATATAGG", "CACACACATATATAGG", "14=2D2="), ("CACACACATATATATAGG", "CACACACATATATATATAGG", "16=2I2="),
             ("AACAACAACAACAAAAA", "AACAACAACAAAAA", "10=3D4="), ("GCACAGCAGTC", "GCACAGTC", "1=2D2=1D5="),
             ("AAAAAAAA", "AAAAAA", "1=1D3=1D2="), ("CAAAGAAAGAAAG", "CAAAGAAAGAAG", "9=1D3="),
             ("CAAAGAAAGAAAG", "CAAAGAAAAGAAAG", "5=1I8="), ("CAAAGAAAGAAAG", "CAAAGAAAAG", "5=4D1I4="),
             ("CAAAGAAAGAAAG", "CAAGAAAG", "1=5D7="), ("CGAAAGAAAGAAAG", "CGAAGAAAG", "2=5D7="),
             ("CGAAAGAAAGAAAC", "CGAAGAAAC", "2=5D7="), ("ATATATATTTTTTAAAGCGCGC", "ATATATATTTTTTAAAGCGCGC", "22=")]
    kats = []
    for rf, sq, cg in cases:
        ir, iq = ref.cig.bases_to_int(rf), ref.cig.bases_to_int(sq)
        ex = ref.cig.expand_cigar(cg)
        row = {"ref": rf, "seq": sq, "cigar": cg}
        for tag, (mb, r) in {"small": (20, 10), "default": (20000, 30)}.items():
            out, sc = ref.aln_sc.align(ir, iq, ex, S, NP, 5, 1, mb, r)
            assert out == ref.aln.align(ir, iq, ex, S, NP, 5, 1, mb, r)
            row[tag] = {"max_b_rows": mb, "r": r, "out": out, "scores": [float(np.float32(x)) for x in sc]}
        kats.append(row)
    json.dump(kats, open(os.path.join(HERE, "align_kats.json"), "w"), indent=0)

    # ---- np_info KATs
    seqs = ["ATATATTTTTTTAAA", "ATATATATATATATATATATTTAA", "ACGATCTCTAGGCAGTTAGCCGAGCAG", "ACCGGCGCAGCAGCAGCAG",
            "TATATATATGCGCGCGGGGATATA", "ATATATATTTTTTAAAGCGCGC", "A" * 105, "ACG" * 40 + "NNN" + "TTTTT"]
    out = []
    for s in seqs:
        info = np.asarray(ref.aln.get_np_info(ref.cig.bases_to_int(s)))
        out.append({"seq": s, "L": info[:, 0, :].T.tolist(), "L_IDX": info[:, 1, :].T.tolist()})
    doc = out[5]  # aln.pyx:182-194 docstring values
    assert doc["L"][0] == [0, 0, 0, 0, 0, 0, 0, 6, 6, 6, 6, 6, 6, 3, 3, 3, 0, 0, 0, 0, 0, 0]
    assert doc["L_IDX"][0] == [0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 0, 1, 2, 0, 0, 0, 0, 0, 0]
    assert doc["L"][1] == [4, 3, 4, 3, 4, 3, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 3, 0, 3, 0]
    assert doc["L_IDX"][1] == [0, 0, 1, 1, 2, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 2, 0]
    json.dump(out, open(os.path.join(HERE, "np_info_kats.json"), "w"))

    # ---- realign_hap KATs (test/test_std_vcf.vcf applied to test_std_ref.fasta)
    contigs = {"chr18": "ACAGCGCTATCAGCAGCTAGCATCAGCATCAG", "chr19": "CAAAGAGGGATTTTCAGCCGCGCAACGAGCAG"}
    haps = [("chr18", 1, "GCAGCGCTATCAGCAGCTAGCATCAGCATCAG", "1X31="),
            ("chr18", 2, "GCACCCTAGCGCTATCAGCAGCTAGCATCAGCATCAG", "1X2=5I29="),
            ("chr19", 1, "CAAAAAGAGAGGGATTTTGAGCCGCGCAACGAGCAG", "1=2I4=2I9=1X17="),
            ("chr19", 2, "CAAAAAGAGGGATTTTGAGCCGCGCAACGAGCAG", "1=2I13=1X17=")]
    hk = []
    for ctg, hap, seq, cg in haps:
        res = ref.bam.realign_hap((ctg, hap, seq, contigs[ctg], ref.cig.expand_cigar(cg)))
        hk.append({"contig": ctg, "hap": hap, "seq": seq, "ref": contigs[ctg], "cigar": cg, "out": res[4]})
    print()
    json.dump(hk, open(os.path.join(HERE, "std_vcf_kats.json"), "w"), indent=0)

    # ---- seeded fuzz with reference outputs
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(20260101)
    fz = []
    for _ in range(400):
        rf, sq, cg, r, mb = synth.fuzz_case(rng, cm)
        ir, iq = ref.cig.bases_to_int(rf), ref.cig.bases_to_int(sq)
        out, sc = ref.aln_sc.align(ir, iq, cg, S, NP, 5, 1, mb, r)
        std = ref_standardized(ref, out, ir, iq)
        fz.append({"ref": rf, "seq": sq, "cigar": cg, "r": r, "max_b_rows": mb, "out": out,
                   "scores": [float(np.float32(x)) for x in sc], "std": ref.cig.collapse_cigar(std)})
    with gzip.open(os.path.join(HERE, "fuzz.json.gz"), "wt") as fh:
        json.dump(fz, fh)
    print("golden vectors written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
