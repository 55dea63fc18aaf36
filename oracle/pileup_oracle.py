"""CPU oracle of the confusion-matrix path  -- TEST INFRASTRUCTURE ONLY.

The reference counts basecaller errors by parsing column 5 of `samtools mpileup -r ctg:start+1-end bam | cut -f5`
(/root/reference/src/bam.pyx:300-314) with a character state machine (bam.pyx:351-510).  samtools is a third-party
binary that is absent from this image and from /root/reference, so its part is RESTATED here from the published
behaviour of samtools/htslib 1.1x (`mpileup.c: pileup_seq`, `sam.c: resolve_cigar2 / bam_plp_insertion`) and is
therefore UNPINNED: nothing in this container can check `mpileup_column5` against a real samtools.  The reference's
own share -- everything after the text exists -- is pinned: `confusion_from_lines` below is checked in
tests/test_oracle.py against the compiled, unmodified `bam.calc_confusion_matrices` fed with the same lines.

mpileup semantics restated (defaults of the reference's command line: no -f, -Q 13, -q 0, --ff UNMAP,SECONDARY,QCFAIL,DUP):
  * one line per reference position of the region that at least one kept read spans (positions nobody spans print
    NO line -- the reference nevertheless advances `pos` by one per LINE, bam.pyx:502; restated as is);
  * per read, in BAM order: ['^' mapq-char at its first position] base-or-'*' ['+'len inserted bases | '-'len N..] ['$'];
    '*' for a position inside a D op, '>' / '<' inside an N op; the indel announcement sits on the LAST position of the
    op before it: next op D (and the current op is not D) => '-', consecutive D ops summed; next op I => '+',
    consecutive I ops summed, also after a D ('*+2AG'); an I that directly follows the clipping or starts the read
    is never announced; deleted bases print as 'N' (no -f);
  * an entry is dropped when the base quality at its query position is below 13 (for '*': the quality of the next
    read base; past the read end: 0); reads stored without qualities (0xff) pass; a line whose entries were all
    dropped reads '*';
  * the reference upper-cases the whole line (bam.pyx:314).
P (padding) ops and the per-position depth cap (-d 8000) are not modelled.
"""
import re

import numpy as np

SKIP_FLAGS = 0x4 | 0x100 | 0x200 | 0x400
_REF_OPS = "MDN=X"
_QRY_OPS = "MIS=X"


class Read:
    """pos: 0-based leftmost reference position; cigar: [(len, op)]; seq: str; qual: bytes / None; flag; mapq."""

    def __init__(self, pos, cigar, seq, qual=None, flag=0, mapq=60):
        self.pos, self.cigar, self.seq, self.qual, self.flag, self.mapq = int(pos), list(cigar), seq, qual, flag, mapq
        self.end = self.pos + sum(n for n, op in self.cigar if op in _REF_OPS)


def _entries(read, min_bq):
    """Yields (ref position, text) for every pileup entry of one read that survives the base-quality filter."""
    g = read.cigar
    lq = len(read.seq)
    x, y = read.pos, 0
    for k, (n, op) in enumerate(g):
        if op in _REF_OPS:
            for t in range(n):
                is_del = op in "DN"
                qpos = y if is_del else y + t
                indel = 0
                if t == n - 1 and k + 1 < len(g):
                    n2, op2 = g[k + 1]
                    if op2 == "D" and op != "D":
                        k2 = k + 1
                        while k2 < len(g) and g[k2][1] == "D":
                            indel -= g[k2][0]; k2 += 1
                    elif op2 == "I":
                        k2 = k + 1
                        while k2 < len(g) and g[k2][1] == "I":
                            indel += g[k2][0]; k2 += 1
                q = 0 if qpos >= lq else (255 if read.qual is None else read.qual[qpos])
                if q >= min_bq:
                    s = ""
                    if x + t == read.pos:
                        s += "^" + chr(min(read.mapq, 93) + 33)
                    if is_del:
                        s += "*" if op == "D" else (">" if not read.flag & 16 else "<")
                    else:
                        c = read.seq[qpos] if qpos < lq else "N"
                        s += ("," if read.flag & 16 else ".") if c == "=" else (c.lower() if read.flag & 16 else c.upper())
                    if indel > 0:
                        q0 = y if is_del else y + n
                        s += f"+{indel}" + read.seq[q0:q0 + indel]
                    elif indel < 0:
                        s += f"{indel}" + "N" * (-indel)
                    if x + t == read.end - 1:
                        s += "$"
                    yield x + t, s
            x += n
        if op in _QRY_OPS:
            y += n


def mpileup_column5(reads, start, end, min_bq=13):
    """Column 5 of `samtools mpileup -r ctg:start+1-end`, one string per printed line, as bam.pyx:310-314 yields them."""
    cols, spanned = {}, set()
    for rd in reads:
        if rd.flag & SKIP_FLAGS:
            continue
        spanned.update(range(max(rd.pos, start), min(rd.end, end)))
        for p, s in _entries(rd, min_bq):
            if start <= p < end:
                cols.setdefault(p, []).append(s)
    return [("".join(cols[p]) if p in cols else "*").upper().strip() for p in sorted(spanned)]


_TOKEN = re.compile(r"\^.|[$*]|[NACGT]|[+-]\d+")
_CODE = {c: i for i, c in enumerate("NACGT")}


def confusion_from_lines(lines, ref, start, end, np_info, max_n=6, max_l=100):
    """Restatement of the line parser of bam.pyx:387-504.  `ref` is the raw contig string (cfg.args.refs[ctg]),
    `np_info` = get_np_info(bases_to_int(ref[start:end+1])) (bam.pyx:381).  Returns subs[5,5], nps, inss, dels (int64)."""
    subs = np.zeros((5, 5), np.int64)
    nps = np.zeros((max_n, max_l + 1, max_l + 1), np.int64)
    inss = np.zeros(max_l + 1, np.int64)
    dels = np.zeros(max_l + 1, np.int64)

    def tracts(pos):        # periods with a tract starting right after this position
        if pos + 1 >= len(np_info):
            return []
        return [(n, int(np_info[pos + 1, 0, n - 1])) for n in range(1, max_n + 1)
                if np_info[pos + 1, 0, n - 1] != 0 and np_info[pos + 1, 1, n - 1] == 0]

    def close(pos, seen_ins, seen_del):        # bam.pyx:409-420, 491-501
        if not seen_ins:
            inss[0] += 1
        if not seen_del:
            dels[0] += 1
        if not seen_ins and not seen_del:
            for n, l in tracts(pos):
                nps[n - 1, l, l] += 1

    for pos, line in enumerate(lines):
        seen_ins = seen_del = True
        rb = _CODE.get(ref[start + pos], 0)
        i = 0
        while i < len(line):
            m = _TOKEN.match(line, i)
            if m is None:                      # bam.pyx:486-489: report and abandon the line
                break
            tok = m.group()
            i = m.end()
            if tok[0] in "^$*":
                continue
            if tok in _CODE:
                subs[rb, _CODE[tok]] += 1
                close(pos, seen_ins, seen_del)
                seen_ins = seen_del = False
                continue
            k = int(tok[1:])
            explained = False
            if tok[0] == "-":
                seen_del = True
                for n, l in tracts(pos):
                    if k % n == 0 and k <= l * n:
                        explained = True
                        nps[n - 1, l, l - k // n] += 1
                    else:
                        nps[n - 1, l, l] += 1
                if not explained:
                    dels[min(max_l, k)] += 1
            else:
                seen_ins = True
                for n, l in tracts(pos):
                    if k % n == 0 and ref[start + pos + 1:start + pos + n + 1] * (k // n) == line[i:i + k]:
                        explained = True
                        nps[n - 1, l, min(max_l, l + k // n)] += 1
                    else:
                        nps[n - 1, l, l] += 1
                if not explained:
                    inss[min(max_l, k)] += 1
            i += k
        close(pos, seen_ins, seen_del)
    return subs, nps, inss, dels


def confusion(reads, ref, start, end, get_np_info, bases_to_int, max_n=6, max_l=100, min_bq=13):
    """calc_confusion_matrices((ctg, start, end)) for reads given as alignments instead of a BAM path."""
    info = get_np_info(bases_to_int(ref[start:end + 1]))
    return confusion_from_lines(mpileup_column5(reads, start, end, min_bq), ref, start, end, info, max_n, max_l)
