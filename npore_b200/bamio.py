"""BAM / FASTA / SAM I/O without pysam -- the callers either side of the hot path.

  read_bam(path)                 BGZF inflate + BAM record decode (SAM spec section 4: `<iiBBHHHiiii` core, 4-bit
                                 `=ACMGRSVTWYHKDBN` bases, `MIDNSHP=XB` op codes)
  get_read_data(bam, fasta, ..)  generator of the 11-tuples of /root/reference/src/bam.pyx:18-47 (same filters: no
                                 secondary / supplementary / unmapped; soft clips stripped from SEQ/QUAL like pysam's
                                 query_alignment_sequence; reference slice = fasta[start:stop].upper(), which is what
                                 get_reference_sequence().upper() yields)
  create_header(...)             bam.pyx:127-145: @HD VN:1.6 SO:coordinate, @SQ per contig, @PG realigner
  write_bam(path, ...)           minimal BGZF/BAM writer (fixtures for the tests; records in input order)
  realign_bam(...)               ingest -> GPU batches -> ordered SAM: what realign.py:75-115 does end to end.  Uses the
                                 native reader / writer of libnpore_b200.so (include/npore_bamio.h: multi-threaded BGZF
                                 inflate, record decode into flat arrays, batch SAM formatting) and never builds
                                 per-read Python objects; the pure-Python functions above remain as the tuple API.
The GPU path is entered through npore_b200.engine (realign_bam) or npore_b200.bam.realign_reads (tuples).
"""
import ctypes as C
import gzip
import io
import os
import struct
import sys
import zlib

import numpy as np

from . import cfg

_SEQ16 = "=ACMGRSVTWYHKDBN"
_OPS = "MIDNSHP=XB"
_SEQ_PAIR = np.frombuffer("".join(a + b for a in _SEQ16 for b in _SEQ16).encode(), dtype=np.uint8).reshape(256, 2)


def read_fasta(path):
    """{contig: sequence} (plain or gzip FASTA)."""
    opener = gzip.open if path.endswith(".gz") else open
    out, name, parts = {}, None, []
    with opener(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    out[name] = "".join(parts)
                name, parts = line[1:].split()[0], []
            else:
                parts.append(line.strip())
    if name is not None:
        out[name] = "".join(parts)
    return out


def _parse_tags(buf):
    tags, p, n = {}, 0, len(buf)
    sizes = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
    fmts = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
    while p + 3 <= n:
        tag, typ = buf[p:p + 2].decode(), chr(buf[p + 2])
        p += 3
        if typ == "A":
            tags[tag] = chr(buf[p]); p += 1
        elif typ in fmts:
            tags[tag] = struct.unpack_from(fmts[typ], buf, p)[0]; p += sizes[typ]
        elif typ in "ZH":
            e = buf.index(b"\0", p)
            tags[tag] = buf[p:e].decode(); p = e + 1
        elif typ == "B":
            sub = chr(buf[p]); cnt = struct.unpack_from("<i", buf, p + 1)[0]
            p += 5
            tags[tag] = list(struct.unpack_from("<" + str(cnt) + fmts[sub][1], buf, p)); p += cnt * sizes[sub]
        else:
            break
    return tags


def read_bam(path):
    """Returns (header_text, [(name, length)], iterator of record dicts).  Whole-file inflate (gzip concatenates the
    BGZF members); fine for the batch sizes realign handles per call."""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"BAM\1":
        raise ValueError(f"{path}: not a BAM file")
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].split(b"\0")[0].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, p)[0]; p += 4
        name = data[p:p + l_name - 1].decode(); p += l_name
        refs.append((name, struct.unpack_from("<i", data, p)[0])); p += 4

    def records(p=p):
        n = len(data)
        while p + 4 <= n:
            block_size = struct.unpack_from("<i", data, p)[0]; p += 4
            ref_id, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, _nref, _npos, _tlen = struct.unpack_from("<iiBBHHHiiii", data, p)
            q = p + 32
            name = data[q:q + l_read_name - 1].decode(); q += l_read_name
            cig = np.frombuffer(data, dtype="<u4", count=n_cigar, offset=q); q += 4 * n_cigar
            packed = np.frombuffer(data, dtype=np.uint8, count=(l_seq + 1) // 2, offset=q); q += (l_seq + 1) // 2
            seq = _SEQ_PAIR[packed].tobytes()[:l_seq].decode("latin-1")
            qual = np.frombuffer(data, dtype=np.uint8, count=l_seq, offset=q); q += l_seq
            tags = _parse_tags(data[q:p + block_size])
            yield {"name": name, "flag": flag, "ref_id": ref_id, "pos": pos, "mapq": mapq, "cigar": cig, "seq": seq,
                   "qual": None if (l_seq and qual[0] == 0xFF) else qual, "tags": tags}
            p += block_size
    return text, refs, records()


def _cigar_string(words):
    return "".join(f"{int(w) >> 4}{_OPS[int(w) & 15]}" for w in words)


def get_read_data(bam_fn, fasta, regions=None, max_reads=0):
    """bam.pyx:18-47.  fasta: path or {contig: seq}.  regions: [(contig, start, stop)] or None for everything (a read
    overlapping two regions is yielded once per region, as pysam's fetch does)."""
    if not os.path.exists(bam_fn):
        print(f"\nERROR: BAM file '{bam_fn}' not found.")
        sys.exit(1)
    fa = read_fasta(fasta) if isinstance(fasta, str) else fasta
    _, refs, recs = read_bam(bam_fn)
    recs = list(recs)
    kept = 0
    for ctg, start, stop in (regions or [(None, 0, 1 << 62)]):
        for r in recs:
            if max_reads and kept >= max_reads:
                return
            if r["flag"] & (0x4 | 0x100 | 0x800) or r["ref_id"] < 0:
                continue
            rname = refs[r["ref_id"]][0]
            ops, lens = r["cigar"] & 15, r["cigar"] >> 4
            ref_span = int(lens[np.isin(ops, (0, 2, 3, 7, 8))].sum())
            rstart, rstop = r["pos"], r["pos"] + ref_span
            if ctg is not None and (rname != ctg or rstop <= start or rstart >= stop):
                continue
            # query_alignment_sequence / qualities: soft-clipped ends removed (hard clips are not in SEQ)
            lead = int(lens[0]) if len(ops) and ops[0] == 4 else (int(lens[1]) if len(ops) > 1 and ops[0] == 5 and ops[1] == 4 else 0)
            trail = int(lens[-1]) if len(ops) and ops[-1] == 4 else (int(lens[-2]) if len(ops) > 1 and ops[-1] == 5 and ops[-2] == 4 else 0)
            seq = r["seq"][lead:len(r["seq"]) - trail]
            quals = "*" if r["qual"] is None else (r["qual"][lead:len(r["qual"]) - trail] + np.uint8(33)).tobytes().decode("latin-1")
            if not quals:
                quals = "*"                                             # bam.pyx:42: an empty quality array is falsy
            hp = r["tags"].get("HP")
            kept += 1
            yield (r["name"], r["flag"], rname, rstart, r["mapq"], _cigar_string(r["cigar"]), rstop, seq.upper(), quals,
                   fa[rname][rstart:rstop].upper(), 0 if hp is None else int(hp))


def create_header(outfile, refs, argv=None):
    """bam.pyx:127-145: (re)creates the SAM with @HD / @SQ / @PG lines.  refs: [(name, length)]."""
    if os.path.dirname(outfile):
        os.makedirs(os.path.dirname(outfile), exist_ok=True)
    with open(outfile, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n")
        for name, length in refs:
            fh.write(f"@SQ\tSN:{name}\tLN:{length}\n")
        fh.write(f"@PG\tID:realigner\tPN:realigner\tVN:{cfg.__version__}\tCL:{' '.join(argv if argv is not None else sys.argv)}\n")


class NativeBam:
    """A BAM file opened through libnpore_b200.so (include/npore_bamio.h): header, per-record columns, flat gathers."""

    COLUMNS = ("ref_id", "pos", "end", "flag", "mapq", "aln_len", "n_cigar", "name_len", "hp", "has_qual")

    def __init__(self, path, n_threads=0, window_bytes=0):
        """window_bytes == 0: the whole file is loaded now.  window_bytes > 0: nothing is loaded yet; every advance() call
        loads the next ~window_bytes of (inflated) records, so files larger than memory stream through."""
        from ._lib import lib
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.npore_bam_open(os.fsencode(path), n_threads, C.byref(h))
        if rc:
            raise (FileNotFoundError if rc == -1 else ValueError)(f"{path}: {self._L.npore_io_last_error().decode()}")
        self._h = h
        t = C.c_char_p()
        self._L.npore_bam_header_text(h, C.byref(t))
        self.text = (t.value or b"").decode()
        self.refs = []
        for i in range(self._L.npore_bam_n_refs(h)):
            name, length = C.c_char_p(), C.c_int64()
            self._L.npore_bam_ref(h, i, C.byref(name), C.byref(length))
            self.refs.append((name.value.decode(), int(length.value)))
        self.window_bytes = int(window_bytes)
        self.n = 0
        for k in self.COLUMNS:
            setattr(self, k, np.zeros(0, np.int32))
        if not self.window_bytes:
            self.advance()

    def advance(self):
        """Load the next window; returns the number of records in it (0 at end of file)."""
        n = int(self._L.npore_bam_advance(self._h, self.window_bytes))
        if n < 0:
            raise ValueError(self._L.npore_io_last_error().decode())
        self.n = n
        cols = {k: np.zeros(max(n, 1), np.int32) for k in self.COLUMNS}
        self._L.npore_bam_columns(self._h, *[cols[k].ctypes.data for k in self.COLUMNS])
        for k, v in cols.items():
            setattr(self, k, v[:n])
        if n and self.window_bytes:          # streaming: the next window inflates in the background while this one is gathered
            self._L.npore_bam_prefetch(self._h, self.window_bytes)
        return n

    def close(self):
        if getattr(self, "_h", None):
            self._L.npore_bam_close(self._h)
            self._h = None

    __del__ = close

    def gather_nib(self, sel, n_threads=0):
        """The selected records' aligned bases as they lie in the file (4 bits per base): (nib uint8, nib_start int64[n]) for
        PackedBatch.from_flat_shared_nib -- no per-base work on the host."""
        sel = np.ascontiguousarray(sel, dtype=np.int64)
        boff = np.zeros(len(sel) + 1, np.int64)
        ps = sel.ctypes.data if len(sel) else None
        if self._L.npore_bam_gather_nib(self._h, len(sel), ps, n_threads, boff.ctypes.data, None, None):
            raise RuntimeError(self._L.npore_io_last_error().decode())
        nib = np.empty(max(int(boff[-1]), 1), np.uint8)
        start = np.zeros(max(len(sel), 1), np.int64)
        if self._L.npore_bam_gather_nib(self._h, len(sel), ps, n_threads, boff.ctypes.data, nib.ctypes.data, start.ctypes.data):
            raise RuntimeError(self._L.npore_io_last_error().decode())
        return nib[:int(boff[-1])], start[:len(sel)]

    def gather_cigar(self, sel, n_threads=0):
        """Only the CIGAR words of the selected records (S / H dropped): (words uint32, offsets int64[n+1])."""
        sel = np.ascontiguousarray(sel, dtype=np.int64)
        off = np.concatenate(([0], np.cumsum(self.n_cigar[sel], dtype=np.int64)))
        cig = np.empty(max(int(off[-1]), 1), np.uint32)
        rc = self._L.npore_bam_gather(self._h, len(sel), sel.ctypes.data if len(sel) else None, n_threads, None, None, None, None,
                                      cig.ctypes.data, off.ctypes.data, None, None)
        if rc:
            raise RuntimeError(self._L.npore_io_last_error().decode())
        return cig[:int(off[-1])], off

    def gather(self, sel, n_threads=0, want_qual=True, want_names=True, want_codes=True, want_cigar=True):
        """Flat arrays of the selected records: dict with seq_ascii, seq_codes, qual_ascii, seq_off, cigar, cig_off, names,
        name_off (soft clips removed, S/H dropped from the CIGAR; bam.pyx:41-44, 59)."""
        sel = np.ascontiguousarray(sel, dtype=np.int64)
        off = lambda col: np.concatenate(([0], np.cumsum(col[sel], dtype=np.int64)))   # noqa: E731
        o = {"seq_off": off(self.aln_len), "cig_off": off(self.n_cigar), "name_off": off(self.name_len)}
        o["seq_ascii"] = np.empty(max(int(o["seq_off"][-1]), 1), np.uint8)
        o["seq_codes"] = np.empty(max(int(o["seq_off"][-1]), 1), np.uint8) if want_codes else None
        o["qual_ascii"] = np.empty(max(int(o["seq_off"][-1]), 1), np.uint8) if want_qual else None
        o["cigar"] = np.empty(max(int(o["cig_off"][-1]), 1), np.uint32) if want_cigar else None
        o["names"] = np.empty(max(int(o["name_off"][-1]), 1), np.uint8) if want_names else None
        ptr = lambda a: None if a is None else a.ctypes.data   # noqa: E731
        rc = self._L.npore_bam_gather(self._h, len(sel), ptr(sel) if len(sel) else None, n_threads, ptr(o["seq_ascii"]), ptr(o["seq_codes"]),
                                      ptr(o["qual_ascii"]), ptr(o["seq_off"]), ptr(o["cigar"]), ptr(o["cig_off"]), ptr(o["names"]), ptr(o["name_off"]))
        if rc:
            raise RuntimeError(self._L.npore_io_last_error().decode())
        return o


def format_sam(bam, sel, g, rle, rle_off, n_threads=0, cols=None, fd=None, offset=0):
    """bam.pyx:83 for the selected records as one bytes-like block (npore_sam_format).  cols: the records' column values
    taken earlier (dict of int32 arrays: flag ref_id pos end mapq has_qual hp) when the reader has moved on since.
    fd: also append the block to that file descriptor at `offset` while formatting (npore_sam_format_fd)."""
    L = bam._L
    names = "".join(n for n, _ in bam.refs).encode()
    rn_off = np.concatenate(([0], np.cumsum([len(n.encode()) for n, _ in bam.refs], dtype=np.int64)))
    rn = np.frombuffer(names, dtype=np.uint8) if names else np.zeros(1, np.uint8)
    rle = np.ascontiguousarray(rle, dtype=np.uint32)
    rle_off = np.ascontiguousarray(rle_off, dtype=np.int64)
    if cols is None:
        cols = take_columns(bam, sel)
    n = len(cols["flag"])
    cap = int(L.npore_sam_bound(n, g["name_off"].ctypes.data, g["seq_off"].ctypes.data, rle_off.ctypes.data, max([len(x) for x, _ in bam.refs] + [1])))
    out = np.empty(max(cap, 1), np.uint8)
    flag, ref_id, pos, end, mapq, hq, hp = (cols[k] for k in ("flag", "ref_id", "pos", "end", "mapq", "has_qual", "hp"))
    qual = g["qual_ascii"] if g["qual_ascii"] is not None else g["seq_ascii"]
    if g["qual_ascii"] is None:
        hq = np.zeros_like(hq)
    args = (n, n_threads, g["names"].ctypes.data, g["name_off"].ctypes.data, flag.ctypes.data, ref_id.ctypes.data,
            rn.ctypes.data, rn_off.ctypes.data, len(bam.refs), pos.ctypes.data, end.ctypes.data, mapq.ctypes.data,
            rle.ctypes.data if len(rle) else None, rle_off.ctypes.data, g["seq_ascii"].ctypes.data, qual.ctypes.data,
            g["seq_off"].ctypes.data, hq.ctypes.data, hp.ctypes.data, out.ctypes.data, cap)
    got = L.npore_sam_format(*args) if fd is None else L.npore_sam_format_fd(*args, int(fd), int(offset))
    if got < 0:
        raise RuntimeError(L.npore_io_last_error().decode())
    return out[:got]


def take_columns(bam, sel):
    return {k: np.ascontiguousarray(getattr(bam, k)[sel], dtype=np.int32) for k in ("flag", "ref_id", "pos", "end", "mapq", "has_qual", "hp")}


def select_reads(bam, regions=None, max_reads=0):
    """Record indices per region like bam.pyx:26-32: fetch(ctg, start, stop) order, no secondary / supplementary /
    unmapped reads, at most max_reads in total.  Yields (contig, indices)."""
    ok = ((bam.flag & (0x4 | 0x100 | 0x800)) == 0) & (bam.ref_id >= 0)
    names = [n for n, _ in bam.refs]
    kept = 0
    for ctg, start, stop in (regions or [(n, 0, l) for n, l in bam.refs]):
        if ctg not in names:
            continue
        sel = np.flatnonzero(ok & (bam.ref_id == names.index(ctg)) & (bam.pos < stop) & (bam.end > start))
        if max_reads:
            sel = sel[:max(0, max_reads - kept)]
        kept += len(sel)
        if len(sel):
            yield ctg, sel


_PIPES = {}


def _pipeline(sub, npt, n_inflight, device=None):
    """Cached PipelinedRealigner per (tables, cfg, device) (same idea as aln._engine)."""
    from .engine import PipelinedRealigner
    dev = int(getattr(cfg.args, "device", 0) or 0) if device is None else int(device)
    key = (np.asarray(sub, np.float32).tobytes(), hash(np.asarray(npt, np.float32).tobytes()), int(cfg.args.max_n), int(cfg.args.max_l), n_inflight)
    for old in [k for k in _PIPES if k[0] != key]:          # tables / cfg changed: drop every cached pipeline
        _PIPES.pop(old).close()
    p = _PIPES.get((key, dev))
    if p is None:
        p = _PIPES[(key, dev)] = PipelinedRealigner(sub, npt, n_inflight=n_inflight, max_n=int(cfg.args.max_n), max_l=int(cfg.args.max_l), device=dev)
    return p


def _read_loads(bam, sel, max_b_rows=20000, r=30):
    """Cell updates per selected record: (Lref + Lseq + n_chunks) * (2r+1), from the record columns alone."""
    ops = (bam.end[sel] - bam.pos[sel]).astype(np.int64) + bam.aln_len[sel]
    return (ops + -(-ops // (max_b_rows - 1))) * (2 * r + 1)


def _realign_segments(bam, segments, fa, pipe, fd, tm, n_threads, max_batch_ops, n_inflight):
    """The three-stage pipeline over an explicit list of (contig, record indices) segments: gather batch k+1 while batch k is
    on the GPU and batch k-1 is formatted and appended to the file descriptor `fd` (not in append mode) -- records in segment
    order.  Returns the records written."""
    import queue
    import threading
    import time
    from .aln import _report
    from .cig import bases_to_int
    from .engine import NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, PackedBatch
    flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
    pending = queue.Queue(maxsize=n_inflight + 1)
    state = {"written": 0, "error": None, "end": os.fstat(fd).st_size}

    def retire_loop():
        while True:
            item = pending.get()
            if item is None:
                return
            if state["error"] is not None:
                continue
            try:
                fut, cols, g, n = item
                t1 = time.perf_counter()
                res, _ = fut.result()
                t2 = time.perf_counter()
                _report(res.status[:n], "realign_read")
                state["end"] += len(format_sam(bam, None, g, res.rle, res.rle_off[:n + 1], n_threads, cols=cols, fd=fd, offset=state["end"]))
                t3 = time.perf_counter()
                tm["gpu_wait"] += t2 - t1; tm["format"] += t3 - t2; tm["write"] += time.perf_counter() - t3
                state["written"] += n
            except Exception as e:                    # noqa: BLE001
                state["error"] = e

    retire = threading.Thread(target=retire_loop, daemon=True)
    retire.start()
    try:
        for ctg, sel in segments:
            contig = fa[ctg]
            ops = np.cumsum((bam.end[sel] - bam.pos[sel]).astype(np.int64) + bam.aln_len[sel])
            cut = 0
            while cut < len(sel) and state["error"] is None:
                stop = max(cut + 1, int(np.searchsorted(ops, (ops[cut - 1] if cut else 0) + max_batch_ops, "right")))
                part = sel[cut:stop]
                t1 = time.perf_counter()
                nib, nib_start = bam.gather_nib(part, n_threads)              # the upload: BAM's own 4-bit bases + CIGAR words
                cig_words, cig_off = bam.gather_cigar(part, n_threads)
                lo, hi = int(bam.pos[part].min()), int(bam.end[part].max())
                packed = PackedBatch.from_flat_shared_nib(bases_to_int(contig[lo:hi].upper()), bam.pos[part].astype(np.int64) - lo, bam.end[part] - bam.pos[part],
                                                          nib, nib_start, bam.aln_len[part], cig_words, cig_off)
                fut = pipe.submit(packed, flags)                              # the GPU starts; the SAM-text gather runs beside it
                g = bam.gather(part, n_threads, want_codes=False, want_cigar=False)      # ASCII bases / qualities / names
                item = (fut, take_columns(bam, part), g, len(part))
                tm["gather"] += time.perf_counter() - t1
                pending.put(item)
                cut = stop
    finally:
        pending.put(None)
        retire.join()
    if state["error"] is not None:
        raise state["error"]
    return state["written"]


def realign_bam_sharded(bam_fn, fasta, devices, out_prefix=None, regions=None, max_reads=0, argv=None, max_batch_ops=64_000_000,
                        n_threads=0, timings=None, n_inflight=2):
    """realign.py:75-115 on several GPUs of one host: the coordinate-sorted reads are cut into len(devices)
    contiguous genomic regions of equal cell-update load (the reference's own unit of distribution is the region,
    util.py:44-93 / bam.pyx:27-28); every region is an independent ingest -> GPU -> SAM-text pipeline on its own device (no
    data-path collective); the host gathers the per-region SAM bodies in region order into ONE file, byte-identical to the
    single-GPU output.  `devices` may name a device more than once (several pipelines on one GPU).  timings: per phase, summed
    over the shards, plus 'shards': [{device, reads, cell_updates, seconds}] and 'concat' seconds."""
    import threading
    import time
    from .bam import _tables
    from .scheduler import cut_balanced
    tm = timings if timings is not None else {}
    for k in ("open", "gather", "gpu_wait", "format", "write", "concat"):
        tm.setdefault(k, 0.0)
    _keep_freed_buffers_mapped()
    t0 = time.perf_counter()
    if out_prefix is not None:
        cfg.args.out_prefix = out_prefix
    if not os.path.exists(bam_fn):
        print(f"\nERROR: BAM file '{bam_fn}' not found.")
        sys.exit(1)
    fa = read_fasta(fasta) if isinstance(fasta, str) else fasta
    bam = NativeBam(bam_fn, n_threads, window_bytes=0)             # region cuts need every record's span
    out = f"{cfg.args.out_prefix}.sam"
    create_header(out, bam.refs, argv)
    sub, npt = _tables()
    G = len(devices)
    segs = list(select_reads(bam, regions, max_reads))
    loads = [_read_loads(bam, sel) for _, sel in segs]
    total = float(sum(int(x.sum()) for x in loads))
    # contiguous cuts of the (contig, start)-ordered read list at equal cumulative load
    shard_segs = [[] for _ in range(G)]
    shard_load = [0] * G
    done = 0.0
    for (ctg, sel), ld in zip(segs, loads):
        owner = cut_balanced(ld, G, done, total)                   # by the read's load midpoint
        for gidx in np.unique(owner):
            m = owner == gidx
            shard_segs[int(gidx)].append((ctg, sel[m]))
            shard_load[int(gidx)] += int(ld[m].sum())
        done += float(ld.sum())
    tm["open"] += time.perf_counter() - t0
    parts = [out if g == 0 else f"{cfg.args.out_prefix}.part{g}.sam" for g in range(G)]
    results = [None] * G
    thread_tm = [{k: 0.0 for k in ("gather", "gpu_wait", "format", "write")} for _ in range(G)]

    def shard_main(g):
        t1 = time.perf_counter()
        try:
            pipe = _pipeline(sub, npt, n_inflight, devices[g])
            # (no O_APPEND: Linux ignores the offset of a pwrite on an append-mode descriptor, and the formatter threads place
            # their slices by offset)
            fd = os.open(parts[g], os.O_WRONLY | (0 if g == 0 else os.O_CREAT | os.O_TRUNC), 0o644)
            try:
                n = _realign_segments(bam, shard_segs[g], fa, pipe, fd, thread_tm[g], max(1, (n_threads or os.cpu_count() or 1) // G),
                                      max_batch_ops, n_inflight)
            finally:
                os.close(fd)
            results[g] = (n, time.perf_counter() - t1, None)
        except Exception as e:      # noqa: BLE001
            results[g] = (0, time.perf_counter() - t1, e)

    with _PIPE_LOCK:                # contexts of different devices are created one after the other
        for g in range(G):
            _pipeline(sub, npt, n_inflight, devices[g])
    threads = [threading.Thread(target=shard_main, args=(g,)) for g in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for n, _, err in results:
        if err is not None:
            raise err
    t2 = time.perf_counter()
    dst = os.open(out, os.O_WRONLY)                                 # host gather: region order = coordinate order
    try:
        end = os.fstat(dst).st_size
        for g in range(1, G):
            with open(parts[g], "rb") as src:
                left = os.fstat(src.fileno()).st_size
                try:                                                # in-kernel copy (no bounce through user space)
                    while left > 0:
                        n = os.copy_file_range(src.fileno(), dst, left, None, end)
                        if n <= 0:
                            break
                        left -= n; end += n
                except (AttributeError, OSError):
                    pass
                while left > 0:                                     # not supported here (or stopped early): plain copy of the rest
                    buf = os.pread(src.fileno(), min(left, 16 << 20), os.fstat(src.fileno()).st_size - left)
                    if not buf:
                        raise OSError(f"short read while gathering {parts[g]}")
                    os.pwrite(dst, buf, end)
                    left -= len(buf); end += len(buf)
            os.remove(parts[g])
    finally:
        os.close(dst)
    tm["concat"] += time.perf_counter() - t2
    for k in ("gather", "gpu_wait", "format", "write"):
        tm[k] += sum(x[k] for x in thread_tm)
    tm["shards"] = [{"device": int(devices[g]), "reads": int(results[g][0]), "cell_updates": int(shard_load[g]), "seconds": round(results[g][1], 4)}
                    for g in range(G)]
    written = sum(n for n, _, _ in results)
    with cfg.counter.get_lock():
        cfg.counter.value += written
    bam.close()
    return written


import threading as _threading  # noqa: E402
_PIPE_LOCK = _threading.Lock()
_MALLOC_TUNED = False


def _keep_freed_buffers_mapped():
    """Once per process: raise glibc's mmap / trim thresholds (mallopt) so that the pipeline's per-window buffers -- nibble and
    CIGAR gathers, SAM text columns, result arrays, the 20 MB SAM blob -- are recycled from the heap instead of being mmap'ed,
    page-faulted by 16 threads at once and munmap'ed again for every window.  Measured on the C2 file (16 host cores): the
    gathers in front of the second and third upload took 8-12 ms instead of 1 ms (page faults + address-space lock beside the
    inflating prefetch threads); with it the host is through all windows of the 3,000-read file after 27 ms and the GPU, not
    the host, limits the call (profiles/r02_ab_experiments.md).  NPORE_NO_MALLOPT=1 leaves the allocator alone."""
    global _MALLOC_TUNED
    if _MALLOC_TUNED or os.environ.get("NPORE_NO_MALLOPT"):
        return
    _MALLOC_TUNED = True
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 32 << 20)            # M_MMAP_THRESHOLD: its maximum (blocks up to 32 MB come from the heap)
        libc.mallopt(-1, 1 << 30)             # M_TRIM_THRESHOLD: keep up to 1 GB of freed heap
    except (OSError, AttributeError):         # not glibc
        pass


def realign_bam(bam_fn, fasta, out_prefix=None, regions=None, max_reads=0, argv=None, max_batch_ops=64_000_000, n_threads=0,
                timings=None, window_bytes=16 << 20, n_inflight=3, devices=None):
    """realign.py:75-115 without pysam / Pool: header, ingest, GPU realignment, records appended in input order
    (= coordinate order for a sorted BAM, which is what the header claims).  Returns the number of records written.
    Flat arrays all the way, as a pipeline: the reader streams the file in windows of ~window_bytes of inflated records
    (native, multi-threaded) and gathers batch k+1 while batch k is on the GPU (n_inflight contexts) and a third thread
    formats (native) and writes batch k-1.  One shared reference slice is uploaded per batch.
    timings: optional dict that receives host seconds per phase (open, gather, gpu_wait, format = formatting + the append to the
    file, write = 0 since the formatter threads write themselves); timings["trace"] = [] collects (seconds, event) of every step.
    devices: CUDA device indices; more than one -> realign_bam_sharded (one region-shard pipeline per device, ordered gather)."""
    if devices is not None and len(devices) > 1:
        return realign_bam_sharded(bam_fn, fasta, list(devices), out_prefix=out_prefix, regions=regions, max_reads=max_reads, argv=argv,
                                   max_batch_ops=max_batch_ops, n_threads=n_threads, timings=timings, n_inflight=n_inflight)
    import queue
    import threading
    import time
    from .bam import _tables
    from .aln import _report
    from .cig import bases_to_int
    from .engine import NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, PackedBatch
    tm = timings if timings is not None else {}
    for k in ("open", "gather", "gpu_wait", "format", "write"):
        tm.setdefault(k, 0.0)
    _keep_freed_buffers_mapped()
    t0 = time.perf_counter()
    trace = tm.get("trace")                       # optional list: (seconds since the call, event) of every pipeline step
    mark = (lambda ev: trace.append((round(time.perf_counter() - t0, 5), ev))) if trace is not None else (lambda ev: None)
    if out_prefix is not None:
        cfg.args.out_prefix = out_prefix
    if not os.path.exists(bam_fn):
        print(f"\nERROR: BAM file '{bam_fn}' not found.")
        sys.exit(1)
    fa = read_fasta(fasta) if isinstance(fasta, str) else fasta
    # several regions are served region by region (bam.pyx:27-28), which needs the whole file at hand
    bam = NativeBam(bam_fn, n_threads, window_bytes=window_bytes if not (regions and len(regions) > 1) else 0)
    mark("bam open")
    streaming = bam.window_bytes > 0
    # the header (and the truncation of a previous output: ~2 ms for a 60 MB file) is written beside the first window's inflate;
    # the writer thread waits for it before it opens the file
    header_err = []

    def write_header():
        try:
            create_header(f"{cfg.args.out_prefix}.sam", bam.refs, argv)
            mark("header written")
        except Exception as e:      # noqa: BLE001
            header_err.append(e)

    header = threading.Thread(target=write_header, daemon=True)
    header.start()
    sub, npt = _tables()
    pipe = _pipeline(sub, npt, n_inflight, devices[0] if devices else None)
    flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
    tm["open"] += time.perf_counter() - t0
    mark("opened")
    # stage 3 (own thread): wait for batch k, format its records, append them to the SAM -- in submit order
    pending = queue.Queue(maxsize=n_inflight + 1)
    state = {"written": 0, "error": None}

    def retire_loop():
        path = f"{cfg.args.out_prefix}.sam"
        header.join()
        if header_err:
            state["error"] = header_err[0]
        fd = os.open(path, os.O_WRONLY) if not header_err else -1
        end = os.path.getsize(path) if fd >= 0 else 0     # the header is there already
        while True:
            item = pending.get()
            if item is None:
                if fd >= 0:
                    os.close(fd)
                return
            if state["error"] is not None:
                continue                              # keep draining so that the producer never blocks
            try:
                fut, cols, g, n = item
                t1 = time.perf_counter()
                res, st_ = fut.result()
                t2 = time.perf_counter()
                mark(f"gpu done n={n} kernels_ms={st_['ms_kernels_total']:.1f} h2d_ms={st_['ms_h2d']:.1f} plan_ms={st_['ms_plan']:.1f} d2h_ms={st_['ms_d2h']:.1f} "
                     f"warps/SM={st_['fwd_warps_per_sm']} context from {1e3 * (st_['wall'][0] - t0):.1f} ms, buffers {1e3 * (st_['wall'][1] - t0):.1f}, "
                     f"to {1e3 * (st_['wall'][2] - t0):.1f} ms")
                _report(res.status[:n], "realign_read")
                end += len(format_sam(bam, None, g, res.rle, res.rle_off[:n + 1], n_threads, cols=cols, fd=fd, offset=end))
                t3 = time.perf_counter()
                mark("formatted + written")
                tm["gpu_wait"] += t2 - t1; tm["format"] += t3 - t2; tm["write"] += time.perf_counter() - t3
                state["written"] += n
                with cfg.counter.get_lock():
                    cfg.counter.value += n
            except Exception as e:                    # noqa: BLE001
                state["error"] = e

    retire = threading.Thread(target=retire_loop, daemon=True)
    retire.start()
    # the text gather runs right after a submit: leave cores to the GPU worker thread that has just been woken (on 16 cores a
    # 16-thread gather delayed the start of the upload by 2 ms) and to the writer
    nt_text = max(1, (n_threads or os.cpu_count() or 1) - int(os.environ.get("NPORE_SPARE_CORES", "3")))
    try:
        kept = 0
        while True:
            if streaming:
                t1 = time.perf_counter()
                more = bam.advance()
                tm["open"] += time.perf_counter() - t1
                mark(f"window of {more} records")
                if not more:
                    break
            for ctg, sel in select_reads(bam, regions, (max_reads - kept) if max_reads else 0):
                kept += len(sel)
                contig = fa[ctg]
                ops = np.cumsum((bam.end[sel] - bam.pos[sel]).astype(np.int64) + bam.aln_len[sel])
                cut = 0
                while cut < len(sel) and state["error"] is None:
                    stop = max(cut + 1, int(np.searchsorted(ops, (ops[cut - 1] if cut else 0) + max_batch_ops, "right")))
                    part = sel[cut:stop]
                    t1 = time.perf_counter()
                    nib, nib_start = bam.gather_nib(part, n_threads)          # the upload: BAM's own 4-bit bases + CIGAR words
                    mark("nib gathered")
                    cig_words, cig_off = bam.gather_cigar(part, n_threads)
                    mark("cigar gathered")
                    lo, hi = int(bam.pos[part].min()), int(bam.end[part].max())
                    # base codes of the batch's reference span only (not the contig: 250 Mb of chr1 for a window that covers 1 Mb)
                    packed = PackedBatch.from_flat_shared_nib(bases_to_int(contig[lo:hi].upper()), bam.pos[part].astype(np.int64) - lo, bam.end[part] - bam.pos[part],
                                                              nib, nib_start, bam.aln_len[part], cig_words, cig_off)
                    fut = pipe.submit(packed, flags)                          # the GPU starts; the SAM-text gather runs beside it
                    mark(f"submitted n={len(part)}")
                    g = bam.gather(part, nt_text, want_codes=False, want_cigar=False)    # ASCII bases / qualities / names
                    mark("text gathered")
                    item = (fut, take_columns(bam, part), g, len(part))
                    tm["gather"] += time.perf_counter() - t1
                    pending.put(item)                                        # blocks while n_inflight + 1 batches are unfinished
                    cut = stop
                if max_reads and kept >= max_reads:
                    break
            if not streaming or (max_reads and kept >= max_reads) or state["error"] is not None:
                break
    finally:
        pending.put(None)
        retire.join()
    if state["error"] is not None:
        raise state["error"]
    written = state["written"]
    bam.close()
    mark("closed")
    return written


# ------------------------------------------------------------------------------------------------ minimal BAM writer
def _bgzf_block(payload: bytes, level: int = 6) -> bytes:
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = comp.compress(payload) + comp.flush()
    bsize = len(cdata) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + cdata +
            struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload)))


def _reg2bin(beg, end):
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


def write_bam(path, header_text, refs, records, level=6):
    """records: dicts with name, flag, ref_id, pos, mapq, cigar (list of (len, op_char)), seq, qual (bytes/None), tags
    ({'HP': int} supported).  level: zlib level of the BGZF members."""
    lut = np.full(256, 15, np.uint8)
    for i, c in enumerate(_SEQ16):
        lut[ord(c)] = i
    out = io.BytesIO()
    text = header_text.encode()
    out.write(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for name, length in refs:
        out.write(struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", length))
    for r in records:
        cig = [(n << 4) | _OPS.index(op) for n, op in r["cigar"]]
        seq = r["seq"]
        codes = np.zeros(len(seq) + (len(seq) & 1), np.uint8)
        codes[:len(seq)] = lut[np.frombuffer(seq.encode("latin-1"), np.uint8)]
        packed = ((codes[0::2] << 4) | codes[1::2]).tobytes()
        qual = bytes([0xFF] * len(seq)) if r.get("qual") is None else bytes(r["qual"])
        tags = b"".join(b"HPi" + struct.pack("<i", v) if k == "HP" else b"" for k, v in r.get("tags", {}).items())
        span = sum(n for n, op in r["cigar"] if op in "MDN=X") or 1
        name = r["name"].encode() + b"\0"
        body = struct.pack("<iiBBHHHiiii", r["ref_id"], r["pos"], len(name), r["mapq"], _reg2bin(r["pos"], r["pos"] + span),
                           len(cig), r["flag"], len(seq), -1, -1, 0)
        body += name + struct.pack(f"<{len(cig)}I", *cig) + packed + qual + tags
        out.write(struct.pack("<i", len(body)) + body)
    raw = out.getvalue()
    with open(path, "wb") as fh:
        for i in range(0, len(raw), 60000):
            fh.write(_bgzf_block(raw[i:i + 60000], level))
        fh.write(_bgzf_block(b""))
