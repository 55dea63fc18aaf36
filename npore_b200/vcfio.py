"""VCF <-> haplotype (+CIGAR) without pysam/tabix -- the steps either side of realign_hap on the standardize_vcf path
(SURVEY.md 8(f) row N3).  Mirrors /root/reference/src/vcf.py (same function names, in-memory records instead of
indexed files):

  read_vcf(path)            plain or gzip VCF -> [Record]
  split_vcf(records)        vcf.py:36-135   phased diploid records -> (hap1 records, hap2 records)
  apply_vcf(records, hap, refs, regions)
                            vcf.py:209-269  haploid records applied to the reference -> [(contig, hap, seq, ref, cigar)]
                            with the expanded CIGAR over '=XID' that align() takes
  gen_vcf(hap_data)         vcf.py:273-380  standardised expanded CIGAR -> haploid records
  merge_vcfs(r1, r2)        vcf.py:139-205  two haploid record lists -> diploid records with GT 1|1, 1|0, 0|1
  standardize_vcf(...)      standardize_vcf.py:10-43 end to end (realign_haps on the GPU in the middle)
Host code only.
"""
import gzip
from collections import namedtuple

from . import cfg

Record = namedtuple("Record", "contig pos ref alts qual gt")      # pos is 1-based like the VCF text; gt e.g. (0, 1)


def read_vcf(path):
    opener = gzip.open if path.endswith(".gz") else open
    out = []
    with opener(path, "rt") as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            f = line.rstrip("\n").split("\t")
            gt = None
            if len(f) > 9 and f[8].split(":")[0] == "GT":
                g = f[9].split(":")[0].replace("/", "|").split("|")
                gt = tuple(int(x) if x != "." else 0 for x in g)
                if len(gt) == 1:
                    gt = (gt[0], gt[0])
            qual = None if f[5] == "." else float(f[5])
            out.append(Record(f[0], int(f[1]), f[3], tuple(f[4].split(",")), qual, gt))
    return out


def split_vcf(records):
    """vcf.py:36-135: per record, which allele (if any) goes on haplotype 1 / 2."""
    h1, h2 = [], []
    for r in records:
        alleles = (r.ref,) + r.alts
        gt = r.gt or (1, 1)
        if len(alleles) == 3:                                   # different variants on the two haplotypes
            if alleles[gt[0]] != "*":
                h1.append(r._replace(alts=(alleles[gt[0]],), gt=None))
            if alleles[gt[1]] != "*":
                h2.append(r._replace(alts=(alleles[gt[1]],), gt=None))
        elif gt[0] and gt[1]:
            h1.append(r._replace(gt=None)); h2.append(r._replace(gt=None))
        elif gt[0]:
            h1.append(r._replace(gt=None))
        elif gt[1]:
            h2.append(r._replace(gt=None))
        elif alleles[0] == alleles[1]:
            pass
        else:
            h1.append(r._replace(gt=None)); h2.append(r._replace(gt=None))
    return h1, h2


def apply_vcf(records, hap, refs, regions=None, min_qual=0):
    """vcf.py:209-269.  refs: {contig: sequence}; regions: [(contig, start, stop)] (default: every contig, whole)."""
    regions = regions or [(c, 0, len(s)) for c, s in refs.items()]
    data = []
    for contig, start, stop in regions:
        ref = refs[contig]
        seq, cig, ref_ptr = [], [], 0
        for r in records:
            if r.contig != contig or not (start <= r.pos - 1 < stop):
                continue
            pos = r.pos - 1
            if (min_qual and not r.qual) or (r.qual and r.qual < min_qual):
                continue
            a0, a1 = r.ref, r.alts[0]
            indel = len(a1) - len(a0)
            if pos < ref_ptr:                                    # overlaps the previous deletion (vcf.py:226-238)
                if indel > 0:
                    seq.append(a1[len(a0):]); cig.append("I" * indel)
                elif indel < 0 and pos == ref_ptr - 1:
                    cig.append("D" * -indel); ref_ptr += -indel
                continue
            seq.append(ref[ref_ptr:pos]); cig.append("=" * (pos - ref_ptr)); ref_ptr = pos
            seq.append(a1)
            for x, y in zip(a0, a1):
                cig.append("=" if x == y else "X"); ref_ptr += 1
            if indel > 0:
                cig.append("I" * indel)
            elif indel < 0:
                cig.append("D" * -indel); ref_ptr += -indel
        cig.append("=" * (len(ref) - ref_ptr)); seq.append(ref[ref_ptr:])
        data.append((contig, hap, "".join(seq), ref, "".join(cig)))
    return data


def gen_vcf(hap_data):
    """vcf.py:273-380: (contig, hap, seq, ref, expanded cigar) -> haploid records (QUAL 60, like the reference)."""
    out = []
    for contig, _hap, seq, ref, cigar in hap_data:
        rp = sp = cp = 0
        n = len(cigar)
        while cp < n:
            op = cigar[cp]
            if op == "=":
                rp += 1; sp += 1; cp += 1
            elif op in "XM":
                if op == "X" or ref[rp] != seq[sp]:
                    out.append(Record(contig, rp + 1, ref[rp], (seq[sp],), 60.0, None))
                rp += 1; sp += 1; cp += 1
            elif op == "D":
                k = 0
                while cp < n and cigar[cp] == "D":
                    k += 1; cp += 1
                if rp > 0:
                    out.append(Record(contig, rp, ref[rp - 1:rp + k], (ref[rp - 1],), 60.0, None))
                rp += k
            elif op == "I":
                k = 0
                while cp < n and cigar[cp] == "I":
                    k += 1; cp += 1
                if rp > 0 and sp > 0:
                    out.append(Record(contig, rp, ref[rp - 1], (ref[rp - 1] + seq[sp:sp + k],), 60.0, None))
                sp += k
            else:
                raise ValueError(f"unrecognized CIGAR operation '{op}'")
    return out


def merge_vcfs(recs1, recs2, contig_order=None):
    """vcf.py:139-205: position-wise merge of two haploid lists (each sorted within a contig)."""
    contigs = contig_order or list(dict.fromkeys([r.contig for r in recs1] + [r.contig for r in recs2]))
    out = []
    for ctg in contigs:
        a = [r for r in recs1 if r.contig == ctg]
        b = [r for r in recs2 if r.contig == ctg]
        i = j = 0
        while i < len(a) or j < len(b):
            p1 = a[i].pos if i < len(a) else float("inf")
            p2 = b[j].pos if j < len(b) else float("inf")
            pos = min(p1, p2)
            h1, h2 = p1 == pos, p2 == pos
            if h1 and h2:
                if (a[i].ref, a[i].alts) == (b[j].ref, b[j].alts):
                    out.append(a[i]._replace(gt=(1, 1)))
                else:
                    out.append(a[i]._replace(gt=(1, 0))); out.append(b[j]._replace(gt=(0, 1)))
            elif h1:
                out.append(a[i]._replace(gt=(1, 0)))
            else:
                out.append(b[j]._replace(gt=(0, 1)))
            i += h1; j += h2
    return out


def write_vcf(path, records, contigs):
    """contigs: [(name, length)]."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as fh:
        fh.write("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n")
        for name, length in contigs:
            fh.write(f"##contig=<ID={name},length={length}>\n")
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write('##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype quality score">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n")
        for r in records:
            gt = "|".join(str(x) for x in r.gt) if r.gt else "."
            q = "." if r.qual is None else f"{r.qual:g}"
            fh.write(f"{r.contig}\t{r.pos}\t.\t{r.ref}\t{','.join(r.alts)}\t{q}\tPASS\t.\tGT\t{gt}\n")


def standardize_vcf(vcf_path, refs, out_path=None, regions=None):
    """standardize_vcf.py:10-43: split -> apply -> realign_haps (GPU) -> gen_vcf -> merge.  refs: {contig: seq} or FASTA path."""
    from .bam import realign_haps
    from .bamio import read_fasta
    if isinstance(refs, str):
        refs = read_fasta(refs)
    refs = {k: v.upper() for k, v in refs.items()}
    h1, h2 = split_vcf(read_vcf(vcf_path))
    hap_data = apply_vcf(h1, 1, refs, regions, getattr(cfg.args, "min_qual", 0)) + apply_vcf(h2, 2, refs, regions, getattr(cfg.args, "min_qual", 0))
    data = realign_haps(hap_data)
    merged = merge_vcfs(gen_vcf([x for x in data if x[1] == 1]), gen_vcf([x for x in data if x[1] == 2]), list(refs))
    if out_path:
        write_vcf(out_path, merged, [(k, len(v)) for k, v in refs.items()])
    return merged
