"""Index construction in the reference's frozen on-disk format, one block.

This is the "-B" (IS_FORCED_ONE_BLOCK) path of ``subread-buildindex`` re-stated as array
operations: sanitise the FASTA (reference src/index-builder.c:789-970), sample a 16-mer every
``gap`` bases from each contig's first base (``build_gene_index`` :150-417), drop 16-mers that
occur more than ``threshold`` times among the sampled ones (``scan_gene_index`` :472-684), hash
them into ``calculate_buckets_by_size`` buckets (src/sorted-hashtable.c:42-75) and sort every
bucket by ``is_1_greater_than_2`` (src/sorted-hashtable.c:1689-1698).  The output files are
byte-identical to the reference builder's (tests/test_index_build.py); the sort runs on the GPU
when one is visible (torch, plumbing only) and in numpy otherwise.

It exists so that benchmarks and tests can make indexes for synthetic genomes without the
reference.  ``build_index_blocks`` is the builder without ``-B``: the index is cut into blocks of
``-M`` megabytes the way ``build_gene_index`` cuts it (:255-330, :353), every block's files
byte-identical to the reference builder's as well.
"""
import gzip
import os

import numpy as np

PADDING = 1210  # MAX_READ_LENGTH, index-builder.c:115
GEHASH_BUCKET_LENGTH = 31  # sorted-hashtable.h:32 (64-bit builds)

_CODE = np.full(256, 3, dtype=np.uint8)  # base2int, subread.h:238
_CODE[: ord("G")] = 2
_CODE[ord("A")] = 0
_CODE[ord("G")] = 1


def calculate_buckets_by_size(expected_size: int, index_gap: int) -> int:
    """sorted-hashtable.c:42-75 for a VER2 table."""
    n = int(expected_size) // GEHASH_BUCKET_LENGTH
    if index_gap >= 3:
        n //= 3
    if n <= 0x3FFFF:
        n = 0x3FFFF + 4
    while any(n % j == 0 for j in range(2, 14)):
        n += 1
    return n


def one_block_bucket_count(index_gap: int) -> int:
    """``-B``: memory_limit 22000 (full) / 11500 (gapped) MB, index-builder.c:1173,1269-1271."""
    memory_limit = 22000 if index_gap == 1 else 11500
    expected = (int(memory_limit * 1024.0 / 8.0) * 1024) & 0xFFFFFFFF
    return calculate_buckets_by_size(expected, index_gap)


def read_fasta_sanitised(path):
    """check_and_convert_FastA (index-builder.c:789-970): upper-case, every non-ACGT character
    (N, '.', '-', IUPAC, ...) becomes 'A'; contigs of <= 16 bases are dropped; the name ends at the
    first blank, tab or '|' is kept (ignore_bar_in_seqnames = 0)."""
    names, seqs, cur, name = [], [], [], None
    opener = gzip.open if str(path).endswith(".gz") else open

    def flush():
        if name is not None:
            s = b"".join(cur)
            if len(s) > 16:
                names.append(name)
                seqs.append(s)

    with opener(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if not line:
                continue
            if line.startswith(b">"):
                flush()
                cur = []
                hdr = line[1:]
                for sep in (b" ", b"\t"):
                    hdr = hdr.split(sep)[0]
                name = hdr.decode()
            else:
                cur.append(line)
        flush()
    table = np.full(256, ord("A"), dtype=np.uint8)
    for c in b"ACGT":
        table[c] = c
        table[c + 32] = c
    return names, [table[np.frombuffer(s, dtype=np.uint8)] for s in seqs]


def _sort_order(bucket, key16, value, parity, use_gpu):
    """Order of (bucket asc, key16 signed asc, value asc if parity == 0 else desc)."""
    tie = np.where(parity == 0, value, np.uint32(0xFFFFFFFF) - value).astype(np.uint32)
    major = (bucket.astype(np.int64) << 16) | (key16.astype(np.int64) + 32768)
    if use_gpu:
        import torch
        dev = torch.device("cuda")
        o1 = torch.argsort(torch.from_numpy(tie.astype(np.int64)).to(dev), stable=True)
        m = torch.from_numpy(major).to(dev)[o1]
        o2 = torch.argsort(m, stable=True)
        return o1[o2].cpu().numpy()
    o1 = np.argsort(tie, kind="stable")
    o2 = np.argsort(major[o1], kind="stable")
    return o1[o2]


def build_index_arrays(names, seqs, gapped=True, threshold=100, buckets=None, use_gpu=None):
    """In-memory form of one index block from sanitised contigs (uint8 arrays of A/C/G/T).
    Returns the SoA arrays ``srcuda_index_upload`` takes plus what the file writer needs."""
    gap = 3 if gapped else 1
    nb = int(buckets) if buckets else one_block_bucket_count(gap)
    if use_gpu is None:
        try:
            import torch
            use_gpu = torch.cuda.is_available()
        except Exception:
            use_gpu = False
    if not seqs:
        raise ValueError("no contig longer than 16 bases")

    keys_l, vals_l, starts, ends = [], [], [], []
    start = PADDING
    for s in seqs:
        L = s.shape[0]
        codes = _CODE[s].astype(np.uint32)
        p = np.arange(0, L - 15, gap, dtype=np.int64)  # window starts while p + 16 <= L
        key = np.zeros(p.shape[0], dtype=np.uint32)
        for i in range(16):
            key |= codes[p + i] << np.uint32(30 - 2 * i)
        keys_l.append(key)
        vals_l.append((p + start).astype(np.uint32))
        starts.append(start)
        ends.append(start + L - 16 + PADDING)        # read_offsets[], index-builder.c:177,205
        start = start + (L - 16) + 2 * PADDING       # next contig's first window, :206
    total_end = starts[-1] + seqs[-1].shape[0]       # linear position after the last base
    if total_end + PADDING > 0xFFFFFFFD:
        raise ValueError("the reference sequences include more than 4 billion bases")
    key = np.concatenate(keys_l)
    val = np.concatenate(vals_l)
    del keys_l, vals_l

    # scan_gene_index: exact occurrence counts of the sampled 16-mers
    srt = np.sort(key)
    edge = np.flatnonzero(np.concatenate(([True], srt[1:] != srt[:-1], [True])))
    run = np.diff(edge)
    heavy = srt[edge[:-1][run > threshold]]
    del srt
    if heavy.shape[0]:
        keep = ~np.isin(key, heavy)
        key, val = key[keep], val[keep]

    bucket = (key % np.uint32(nb)).astype(np.uint32)
    key16 = (key // np.uint32(nb)).astype(np.uint16).view(np.int16)
    parity = ((key % np.uint32(791)) % np.uint32(2)).astype(np.uint8)
    order = _sort_order(bucket, key16, val, parity, use_gpu)
    bucket, key16, val = bucket[order], key16[order], val[order]
    counts = np.bincount(bucket, minlength=nb).astype(np.int64)
    bucket_start = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(counts, out=bucket_start[1:])

    # 2-bit genome, gvindex_set / gvindex_dump (gene-value-index.c:135-187): start_point 0
    length = total_end + PADDING
    n_bytes = (length >> 2) + 1
    bases = np.zeros(n_bytes * 4, dtype=np.uint8)
    for s, st in zip(seqs, starts):
        bases[st:st + s.shape[0]] = _CODE[s]
    b4 = bases.reshape(-1, 4)
    packed = (b4[:, 0] | (b4[:, 1] << 2) | (b4[:, 2] << 4) | (b4[:, 3] << 6)).astype(np.uint8)
    return dict(index_gap=gap, padding=PADDING, buckets_number=nb, n_items=int(key16.shape[0]),
                bucket_start=bucket_start.astype(np.uint32), keys=key16, vals=val.astype(np.uint32), item_bucket=bucket,
                start_point=0, length=int(length), values=packed, read_offsets=np.array(ends, dtype=np.uint32),
                names=list(names), contig_starts=np.array(starts, dtype=np.int64), removed_keys=int(heavy.shape[0]))


def write_index_files(prefix, a, buckets_per_piece=1 << 20, block_no=0, write_reads=True):
    """<prefix>.NN.b.tab (gehash_dump, sorted-hashtable.c:1700-1908), .NN.b.array and .reads.  The table is written in
    pieces of ``buckets_per_piece`` buckets so that a human-sized block (6.3 GB) needs no second copy in memory."""
    nb, n_items = int(a["buckets_number"]), int(a["n_items"])
    starts_items = a["bucket_start"].astype(np.int64)
    counts = np.diff(starts_items)
    head = bytearray(b"2subindx")
    head += np.array([0x0102, 2, a["padding"], 0x0101, 2, a["index_gap"], 0], dtype="<i2").tobytes()
    head += np.array([n_items], dtype="<i8").tobytes() + np.array([nb], dtype="<i4").tobytes()
    keys, vals = a["keys"], a["vals"]
    with open(prefix + ".%02d.b.tab" % block_no, "wb") as fh:
        fh.write(bytes(head))
        for b0 in range(0, nb, buckets_per_piece):
            b1 = min(nb, b0 + buckets_per_piece)
            i0, i1 = int(starts_items[b0]), int(starts_items[b1])
            cnt = counts[b0:b1]
            body = np.zeros(8 * (b1 - b0) + 6 * (i1 - i0), dtype=np.uint8)
            hdr_at = 8 * np.arange(b1 - b0, dtype=np.int64) + 6 * (starts_items[b0:b1] - i0)
            cnt_bytes = cnt.astype("<i4").view(np.uint8).reshape(-1, 4)
            for k in range(4):
                body[hdr_at + k] = cnt_bytes[:, k]
                body[hdr_at + 4 + k] = cnt_bytes[:, k]  # space_size == current_items (sorted-hashtable.c:267)
            if i1 > i0:
                item_bucket = np.repeat(np.arange(b1 - b0, dtype=np.int64), cnt)
                within = np.arange(i1 - i0, dtype=np.int64) - (starts_items[b0:b1] - i0)[item_bucket]
                key_at = hdr_at[item_bucket] + 8 + 2 * within
                val_at = hdr_at[item_bucket] + 8 + 2 * cnt[item_bucket] + 4 * within
                kb = np.ascontiguousarray(keys[i0:i1]).astype("<i2").view(np.uint8).reshape(-1, 2)
                vb = np.ascontiguousarray(vals[i0:i1]).astype("<u4").view(np.uint8).reshape(-1, 4)
                for k in range(2):
                    body[key_at + k] = kb[:, k]
                for k in range(4):
                    body[val_at + k] = vb[:, k]
            fh.write(body.tobytes())
        fh.write(b"\0")  # trailing byte: is_small_table = 0
    with open(prefix + ".%02d.b.array" % block_no, "wb") as fh:
        fh.write(np.array([a["start_point"], a["length"]], dtype="<u4").tobytes())
        fh.write(a["values"].tobytes())
    if not write_reads:
        return
    with open(prefix + ".reads", "wb") as fh:  # index-builder.c:189-194
        for e, nm in zip(a["read_offsets"], a["names"]):
            fh.write(b"%d\t%s\n" % (int(e), nm.encode()))


def build_index(fasta_path, prefix, gapped=True, threshold=100, buckets=None, use_gpu=None):
    """FASTA -> <prefix>.00.b.tab, <prefix>.00.b.array and <prefix>.reads.  Returns a summary dict."""
    names, seqs = read_fasta_sanitised(fasta_path)
    a = build_index_arrays(names, seqs, gapped, threshold, buckets, use_gpu)
    write_index_files(prefix, a)
    return dict(buckets=a["buckets_number"], items=a["n_items"], gap=a["index_gap"], removed_keys=a["removed_keys"],
                length=a["length"], contigs=len(seqs))


MIN_READ_SPLICING = 2000000  # index-builder.c:41


def expected_items_for_memory(memory_mb: int) -> int:
    """``-M``: items per block, index-builder.c:1269 (8000 MB by default, :1026)."""
    return (int(memory_mb * 1024.0 / 8.0) * 1024) & 0xFFFFFFFF


def plan_blocks(read_len, inserted, first_window_of_contig, contig_of_window, expected_items, gap):
    """Where ``build_gene_index`` cuts the index (index-builder.c:255-330, :353), as ranges of sampled windows.

    The builder walks the windows in order; behind every window it looks at the number of items the current table holds and
    at how far into its contig it is (``read_len``): a table that has reached ``expected_items`` is closed as soon as the
    walk is more than 2 Mbp into a contig or within the first 32 bases of one.  It then steps BACK -- to the contig's first
    window in the second case, by 1 999 990 bases (+16) in the first -- so consecutive blocks overlap.  Returns
    [(first_window, last_window)] inclusive."""
    n = read_len.shape[0]
    cum = np.concatenate(([0], np.cumsum(inserted, dtype=np.int64)))
    may_cut = (read_len > MIN_READ_SPLICING) | (read_len < 32)
    cut_at = np.flatnonzero(may_cut)
    back = MIN_READ_SPLICING - 10
    back = back - back % 3 + 1                       # :269-271
    blocks, first = [], 0
    while True:
        # first window j >= first with cum[j + 1] - cum[first] >= expected_items and may_cut[j]
        need = cum[first] + expected_items
        j_items = int(np.searchsorted(cum, need, side="left")) - 1       # smallest j with cum[j + 1] >= need
        j_items = max(j_items, first)
        k = int(np.searchsorted(cut_at, j_items, side="left"))
        if j_items >= n or k >= cut_at.shape[0]:
            blocks.append((first, n - 1))
            return blocks
        j = int(cut_at[k])
        blocks.append((first, j))
        if read_len[j] < MIN_READ_SPLICING:
            first = int(first_window_of_contig[contig_of_window[j]])
        else:
            if (back - 16) % gap:
                raise ValueError("the step back does not land on a sampled window")   # 1 999 974 is a multiple of 3 (and of 1)
            first = j - (back - 16) // gap
        if len(blocks) > 99:
            raise ValueError("more than 99 index blocks")


def build_index_blocks(names, seqs, gapped=True, threshold=100, memory_mb=8000, use_gpu=False):
    """The blocks of an index built WITHOUT ``-B`` (subread-buildindex -M memory_mb [-F]): a list of the dicts
    ``build_index_arrays`` returns, one per block, each with its own table, its own stretch of the 2-bit genome
    (``start_point`` = linear position of its first window) and the shared contig table."""
    gap = 3 if gapped else 1
    expected = expected_items_for_memory(memory_mb)
    nb = calculate_buckets_by_size(expected, gap)
    if not seqs:
        raise ValueError("no contig longer than 16 bases")
    keys_l, vals_l, rl_l, cw_l, first_w, starts, ends = [], [], [], [], [], [], []
    start, n_w = PADDING, 0
    for c, s in enumerate(seqs):
        L = s.shape[0]
        codes = _CODE[s].astype(np.uint32)
        p = np.arange(0, L - 15, gap, dtype=np.int64)
        key = np.zeros(p.shape[0], dtype=np.uint32)
        for i in range(16):
            key |= codes[p + i] << np.uint32(30 - 2 * i)
        keys_l.append(key)
        vals_l.append((p + start).astype(np.uint32))
        rl_l.append(p + 16)
        cw_l.append(np.full(p.shape[0], c, dtype=np.int32))
        first_w.append(n_w)
        n_w += p.shape[0]
        starts.append(start)
        ends.append(start + L - 16 + PADDING)
        start = start + (L - 16) + 2 * PADDING
    total_end = starts[-1] + seqs[-1].shape[0]
    if total_end + PADDING > 0xFFFFFFFD:
        raise ValueError("the reference sequences include more than 4 billion bases")
    key, val = np.concatenate(keys_l), np.concatenate(vals_l)
    read_len, contig_of_window = np.concatenate(rl_l), np.concatenate(cw_l)
    del keys_l, vals_l, rl_l, cw_l
    # scan_gene_index runs once over the whole genome: a 16-mer sampled more than `threshold` times is in no block
    srt = np.sort(key)
    edge = np.flatnonzero(np.concatenate(([True], srt[1:] != srt[:-1], [True])))
    heavy = srt[edge[:-1][np.diff(edge) > threshold]]
    del srt
    inserted = ~np.isin(key, heavy) if heavy.shape[0] else np.ones(key.shape[0], dtype=bool)
    # the whole 2-bit genome once; a block's array is a stretch of it with the bases outside the block cleared
    n_bases = ((total_end + PADDING) >> 2) * 4 + 8
    bases = np.zeros(n_bases, dtype=np.uint8)
    for s, st in zip(seqs, starts):
        bases[st:st + s.shape[0]] = _CODE[s]

    out = []
    for first, last in plan_blocks(read_len, inserted, np.array(first_w, dtype=np.int64), contig_of_window, expected, gap):
        sel = np.flatnonzero(inserted[first:last + 1]) + first
        k, v = key[sel], val[sel]
        bucket = (k % np.uint32(nb)).astype(np.uint32)
        key16 = (k // np.uint32(nb)).astype(np.uint16).view(np.int16)
        parity = ((k % np.uint32(791)) % np.uint32(2)).astype(np.uint8)
        order = _sort_order(bucket, key16, v, parity, use_gpu)
        bucket, key16, v = bucket[order], key16[order], v[order]
        bucket_start = np.zeros(nb + 1, dtype=np.int64)
        np.cumsum(np.bincount(bucket, minlength=nb).astype(np.int64), out=bucket_start[1:])
        # gvindex_init(offset of the block's first window) / gvindex_set up to the last window (or the contig's last base when
        # the block ends with the input) / gvindex_dump (gene-value-index.c:29-41,135-187)
        start_point = 0 if not out else int(val[first])
        last_base = total_end if last == key.shape[0] - 1 else int(val[last]) + 16
        length = last_base - start_point + PADDING
        base0 = start_point - start_point % 4
        n_bytes = ((length + start_point - base0) >> 2) + 1
        piece = np.zeros(n_bytes * 4, dtype=np.uint8)
        lo, hi = start_point, min(last_base, base0 + n_bytes * 4)
        piece[lo - base0:hi - base0] = bases[lo:hi]
        b4 = piece.reshape(-1, 4)
        packed = (b4[:, 0] | (b4[:, 1] << 2) | (b4[:, 2] << 4) | (b4[:, 3] << 6)).astype(np.uint8)
        out.append(dict(index_gap=gap, padding=PADDING, buckets_number=nb, n_items=int(key16.shape[0]),
                        bucket_start=bucket_start.astype(np.uint32), keys=key16, vals=v.astype(np.uint32), item_bucket=bucket,
                        start_point=start_point, length=int(length), values=packed, read_offsets=np.array(ends, dtype=np.uint32),
                        names=list(names), contig_starts=np.array(starts, dtype=np.int64), removed_keys=int(heavy.shape[0]),
                        windows=(int(first), int(last))))
    return out


def build_index_multiblock(fasta_path, prefix, gapped=True, threshold=100, memory_mb=8000, use_gpu=False):
    """FASTA -> <prefix>.NN.b.tab / .NN.b.array for every block and <prefix>.reads: ``subread-buildindex -M memory_mb [-F]``."""
    names, seqs = read_fasta_sanitised(fasta_path)
    blocks = build_index_blocks(names, seqs, gapped, threshold, memory_mb, use_gpu)
    for n, a in enumerate(blocks):
        write_index_files(prefix, a, block_no=n, write_reads=(n == len(blocks) - 1))
    return dict(blocks=len(blocks), buckets=blocks[0]["buckets_number"], items=[a["n_items"] for a in blocks], gap=blocks[0]["index_gap"],
                contigs=len(seqs))


def build_index_arrays_device(names, codes, contig_lens, gapped=True, threshold=100, buckets=None, device=None,
                              chunk=1 << 26):
    """The same one-block index as ``build_index_arrays`` for genomes of human size, built with torch on one
    GPU (plumbing): ``codes`` is a uint8 tensor on the device holding the 2-bit codes (A0 G1 C2 T3, base2int) of all
    contigs back to back, ``contig_lens`` their lengths.  Peak device memory is about 45 bytes per sampled 16-mer
    (a 3.1 Gbp gapped index: 1.03 G items, ~50 GB).  Returns the same dict (numpy arrays on the host) without
    ``item_bucket``; ``tests/test_index_build.py`` holds it to the file writer's byte-identical output on a small genome.

    Order inside a bucket (sorted-hashtable.c:1689-1698): key16 ascending, ties by position ascending when
    ``(key % 791) % 2 == 0`` and descending otherwise.  Windows are enumerated in ascending position, so one stable sort
    by (bucket, key16) gives the ascending ties; listing the odd-parity windows in descending position first gives theirs."""
    import torch
    dev = codes.device if device is None else torch.device(device)
    gap = 3 if gapped else 1
    nb = int(buckets) if buckets else one_block_bucket_count(gap)
    lens = [int(x) for x in contig_lens]
    if not lens or min(lens) <= 16:
        raise ValueError("every contig must be longer than 16 bases")
    starts, ends = [], []
    start = PADDING
    for L in lens:
        starts.append(start)
        ends.append(start + L - 16 + PADDING)
        start = start + (L - 16) + 2 * PADDING
    total_end = starts[-1] + lens[-1]
    if total_end + PADDING > 0xFFFFFFFD:
        raise ValueError("the reference sequences include more than 4 billion bases")
    n_win = [(L - 15 + gap - 1) // gap for L in lens]
    win_first = np.concatenate(([0], np.cumsum(n_win))).astype(np.int64)
    total = int(win_first[-1])
    cat_first = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    t_win_first = torch.from_numpy(win_first).to(dev)
    t_cat_first = torch.from_numpy(cat_first[:-1].copy()).to(dev)
    t_lin_first = torch.tensor(starts, dtype=torch.int64, device=dev)

    major = torch.empty(total, dtype=torch.int64, device=dev)      # bucket << 16 | key16
    value = torch.empty(total, dtype=torch.int64, device=dev)      # linear position
    odd = torch.empty(total, dtype=torch.bool, device=dev)
    for lo in range(0, total, chunk):
        hi = min(total, lo + chunk)
        w = torch.arange(lo, hi, device=dev, dtype=torch.int64)
        c = torch.searchsorted(t_win_first, w, right=True) - 1
        p = (w - t_win_first[c]) * gap
        g = t_cat_first[c] + p
        key = torch.zeros(hi - lo, dtype=torch.int64, device=dev)
        for i in range(16):
            key |= codes[g + i].to(torch.int64) << (30 - 2 * i)
        major[lo:hi] = ((key % nb) << 16) | (key // nb)
        value[lo:hi] = t_lin_first[c] + p
        odd[lo:hi] = ((key % 791) % 2) == 1
        del w, c, p, g, key
    order0 = torch.cat((torch.nonzero(~odd).flatten(), torch.nonzero(odd).flatten().flip(0)))
    del odd
    m0 = major[order0]
    del major
    m_sorted, o2 = torch.sort(m0, stable=True)
    del m0
    order = order0[o2]
    del order0, o2
    val = value[order]
    del value, order

    # scan_gene_index: 16-mers sampled more than `threshold` times are left out altogether
    removed = 0
    _, counts = torch.unique_consecutive(m_sorted, return_counts=True)
    heavy = counts > threshold
    if bool(heavy.any()):
        removed = int(heavy.sum().item())
        keep = ~torch.repeat_interleave(heavy, counts)
        m_sorted, val = m_sorted[keep], val[keep]
    del counts, heavy

    bucket = m_sorted >> 16
    key16 = (m_sorted & 0xFFFF).to(torch.int16)
    del m_sorted
    bcount = torch.bincount(bucket, minlength=nb)
    del bucket
    bucket_start = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    torch.cumsum(bcount, 0, out=bucket_start[1:])

    length = total_end + PADDING
    n_bytes = (length >> 2) + 1
    bases = torch.zeros(n_bytes * 4, dtype=torch.uint8, device=dev)
    for ci, (st, L) in enumerate(zip(starts, lens)):
        bases[st:st + L] = codes[int(cat_first[ci]):int(cat_first[ci]) + L]
    b4 = bases.view(-1, 4)
    packed = b4[:, 0] | (b4[:, 1] << 2) | (b4[:, 2] << 4) | (b4[:, 3] << 6)
    del bases, b4

    out = dict(index_gap=gap, padding=PADDING, buckets_number=nb, n_items=int(key16.shape[0]),
               bucket_start=bucket_start.cpu().numpy().astype(np.uint32), keys=key16.cpu().numpy(),
               vals=val.cpu().numpy().astype(np.uint32), start_point=0, length=int(length),
               values=packed.cpu().numpy(), read_offsets=np.array(ends, dtype=np.uint32), names=list(names),
               contig_starts=np.array(starts, dtype=np.int64), removed_keys=removed)
    return out
