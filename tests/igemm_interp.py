"""CPU interpreter of the implicit-GEMM descriptors (test infrastructure, not product code).

Executes exactly the gather semantics the sm_100a kernels implement -- TMA boxes with zero fill outside the tensor
bounds, tap tables, parity views, output views, split/scatter of wgrad tiles -- with numpy on fake base addresses,
so that the planners in csrc/conv_plan.cu can be checked against torch.nn.functional on a machine without a GPU.
"""
import numpy as np

FAKE_BASE = 0x10000000  # fake "device address" of element 0 of a buffer


def _tmap_gather(tm, buf, coords_start, elem_size=2):
    """Emulate one TMA box load: returns an array shaped box[::-1] (slowest dim first)."""
    rank = tm.rank
    dims = [int(tm.dims[i]) for i in range(rank)]
    strides = [elem_size] + [int(tm.strides[i]) for i in range(rank - 1)]
    box = [int(tm.box[i]) for i in range(rank)]
    base_elem = (int(tm.base or 0) - FAKE_BASE) // elem_size
    assert (int(tm.base or 0) - FAKE_BASE) % 16 == 0, "TMA base must be 16-byte aligned"
    for s in strides[1:]:
        assert s % 16 == 0, "TMA strides must be multiples of 16 bytes"
    assert box[0] * elem_size == tm.swizzle_bytes
    idx = np.zeros([1] * rank, dtype=np.int64)
    valid = np.ones([1] * rank, dtype=bool)
    for d in range(rank):
        c = coords_start[d] + np.arange(box[d], dtype=np.int64)
        shape = [1] * rank
        shape[rank - 1 - d] = box[d]
        ok = (c >= 0) & (c < dims[d])
        idx = idx + (np.where(ok, c, 0) * (strides[d] // elem_size)).reshape(shape)
        valid = valid & ok.reshape(shape)
    flat = base_elem + idx
    out = np.where(valid, buf[np.clip(flat, 0, buf.size - 1)], 0.0)
    assert (flat[valid] < buf.size).all() and (flat[valid] >= 0).all(), "in-bounds coordinates left the buffer"
    return out


def run_fprop(desc, abuf, bbuf, outbuf, bias=None):
    """Interpret an fpg_igemm_fprop_desc. abuf/bbuf/outbuf are flat float arrays standing for the device buffers."""
    cblk, bn = desc.cblk, desc.block_n
    cpt = desc.c_per_tap // cblk
    sub_per_stage = 64 // cblk
    assert desc.num_sub % sub_per_stage == 0
    assert desc.tile_h * desc.tile_w in (128, 256)
    ktot = int(desc.b.dims[0])
    b_base = (int(desc.b.base or 0) - FAKE_BASE) // 2
    n_total = desc.n_blocks * bn
    bmat = bbuf[b_base:b_base + n_total * ktot].reshape(n_total, ktot)
    o = desc.out
    o_base = (int(o.base or 0) - FAKE_BASE) // (4 if o.fp32 else 2)
    pair = 2 if desc.cta_pair else 1
    if desc.cta_pair:  # each CTA of the pair loads half of the weight rows of an n-block
        assert int(desc.b.box[1]) * 2 == bn and desc.tile_h * desc.tile_w == 128
    else:
        assert int(desc.b.box[1]) == bn
    regions = [(desc.a, desc.tiles_y * pair, desc.tiles_x, desc.tile_h, desc.tile_w, 0)]
    if desc.tiles_x1 > 0:  # second tile region: edge columns with their own tile shape
        assert not desc.cta_pair and desc.tile_h1 * desc.tile_w1 == 128 and desc.tiles_x * desc.tile_w == desc.x_org1
        regions.append((desc.a1, desc.tiles_y1, desc.tiles_x1, desc.tile_h1, desc.tile_w1, desc.x_org1))
    for n in range(desc.n_img):
      for (amap, n_ty, n_tx, tile_h, tile_w, x_org) in regions:
        for ty in range(n_ty):  # pair tiles: rank r of the cluster takes row-tile 2*ty + r
            for tx in range(n_tx):
                x0, y0 = x_org + tx * tile_w, ty * tile_h
                acc = np.zeros((tile_h, tile_w, n_total), dtype=np.float64)
                for sub in range(desc.num_sub):
                    tap, chunk = divmod(sub, cpt)
                    t = desc.taps[tap]
                    a = _tmap_gather(amap, abuf, [t.c0 + chunk * cblk, x0 + t.dx, t.plane, y0 + t.dy, n])
                    a = a.reshape(tile_h, tile_w, cblk)  # (n=1, y, plane=1, x, c)
                    acc += a.astype(np.float64) @ bmat[:, sub * cblk:(sub + 1) * cblk].T.astype(np.float64)
                if bias is not None:
                    acc += bias[None, None, :n_total]
                if desc.act == 1:
                    acc = np.maximum(acc, 0)
                elif desc.act == 2:
                    acc = np.where(acc > 0, acc, 0.2 * acc)
                elif desc.act == 3:
                    acc = np.tanh(acc)
                for ry in range(tile_h):
                    py = y0 + ry
                    if py >= o.valid_h:
                        continue
                    for rx in range(tile_w):
                        px = x0 + rx
                        if px >= o.valid_w:
                            continue
                        off = (o_base + n * o.stride_n + (py * o.mul_y + o.off_y) * o.stride_y +
                               (px * o.mul_x + o.off_x) * o.stride_x)
                        outbuf[off:off + n_total] = acc[ry, rx]


def run_rows(desc, abuf, bbuf, outbuf, bias=None):
    """Interpret an fpg_igemm_rows_desc (row-stationary kernel): one (128 + cols - 1)-pixel box per input row of the
    patch; accumulator `a` of a tile uses input row j as filter row j - a and the column taps as shifted windows of
    that box."""
    cblk, bn = desc.cblk, desc.block_n
    R, S, TH = desc.rows, desc.cols, desc.tile_rows
    assert 2 * TH * bn <= 512 and int(desc.a.box[1]) == 128 + S - 1 and int(desc.a.box[0]) == cblk
    assert int(desc.b.box[0]) == cblk and int(desc.b.box[1]) == bn
    assert desc.b_stages == R or desc.b_stages >= TH + 1
    smem = (desc.a_stages * (((128 + S - 1) * cblk * 2 + 1023) // 1024 * 1024) + desc.b_stages * S * bn * cblk * 2)
    assert smem <= 224 * 1024, smem
    ktot = int(desc.b.dims[0])
    b_base = (int(desc.b.base or 0) - FAKE_BASE) // 2
    bmat = bbuf[b_base:b_base + bn * ktot].reshape(bn, ktot)
    o = desc.out
    o_base = (int(o.base or 0) - FAKE_BASE) // (4 if o.fp32 else 2)
    for n in range(desc.n_img):
        for ty in range(desc.tiles_y):
            for tx in range(desc.tiles_x):
                x0, y0 = tx * 128 + desc.dx0, ty * TH + desc.dy0
                acc = np.zeros((TH, 128, bn), dtype=np.float64)
                for j in range(TH + R - 1):
                    row = _tmap_gather(desc.a, abuf, [0, x0, 0, y0 + j, n]).reshape(128 + S - 1, cblk)
                    for a in range(TH):
                        r = j - a
                        if r < 0 or r >= R:
                            continue
                        for s in range(S):
                            t = desc.tap_of[r * S + s]
                            acc[a] += row[s:s + 128].astype(np.float64) @ bmat[:, t * cblk:(t + 1) * cblk].T
                if bias is not None:
                    acc += bias[None, None, :bn]
                if desc.act == 1:
                    acc = np.maximum(acc, 0)
                elif desc.act == 2:
                    acc = np.where(acc > 0, acc, 0.2 * acc)
                elif desc.act == 3:
                    acc = np.tanh(acc)
                for a in range(TH):
                    py = ty * TH + a
                    if py >= o.valid_h:
                        continue
                    for rx in range(128):
                        px = tx * 128 + rx
                        if px >= o.valid_w:
                            continue
                        off = (o_base + n * o.stride_n + (py * o.mul_y + o.off_y) * o.stride_y +
                               (px * o.mul_x + o.off_x) * o.stride_x)
                        outbuf[off:off + bn] = acc[a, rx]


def run_wgrad(desc, xbuf, ybuf, dw, stride_k, stride_c, k_valid, c_valid):
    """Interpret an fpg_igemm_wgrad_desc including shifted operands (shift atoms / shift groups), the split
    reduction and the scatter into dw (flat fp array)."""
    assert desc.tile_h * desc.tile_w == 64
    if desc.cta_pair:
        return _run_wgrad_pair(desc, xbuf, ybuf, dw, stride_k, stride_c, k_valid, c_valid)
    M, N = desc.x_atoms * desc.x_ca, desc.y_atoms * desc.y_ca
    YSH, SETS = max(desc.y_shifts, 1), max(desc.y_sets, 1)
    YS = YSH * SETS  # MMA groups per stage, each with its own accumulator columns
    assert M in (64, 128, 256) and N % 16 == 0 and 16 <= N <= 256 and (2 if M > 128 else 1) * YS * N <= 512
    if desc.x_shift_atoms or desc.y_shift_atoms or YS > 1:
        assert desc.tile_h == 1
    assert not (desc.y_shift_atoms and YSH > 1) and not (desc.x_shift_atoms and M > 128)
    assert SETS == 1 or desc.y_shift_atoms
    NX = desc.x_groups if desc.x_taps_mode else desc.x_groups * desc.x_ntaps
    NY = -(-desc.y_groups // SETS) if desc.y_taps_mode else desc.y_groups * desc.y_ntaps
    total_kt = desc.n_img * desc.kt_y * desc.kt_x
    taps_total = desc.taps_r * desc.taps_s

    def operand_tile(tm, buf, taps, taps_mode, ntaps, groups, atoms, ca, shift_atoms, shifts, idx, n, y0, x0):
        """returns [shift group] -> (64 x atoms*ca matrix), and per-column (tap index or -1, channel)"""
        extent = 64 + (atoms - 1 if shift_atoms else 0) + (shifts - 1)
        assert int(tm.box[1]) * int(tm.box[3]) == extent and int(tm.box[0]) == ca
        meta = []
        boxes = []
        for a in range(1 if shift_atoms else atoms):
            if taps_mode:
                tap = idx * atoms + a
                dummy = tap >= ntaps
                if dummy:
                    tap = 0
                coff = 0
            else:
                tap = idx // groups
                coff = ((idx % groups) * atoms + a) * ca
            t = taps[tap]
            boxes.append(_tmap_gather(tm, buf, [t.c0 + coff, x0 + t.dx, t.plane, y0 + t.dy, n]).reshape(extent, ca))
        for a in range(atoms):
            if taps_mode:
                tap = idx * atoms + a
                coff = 0
            else:
                tap = idx // groups
                coff = ((idx % groups) * atoms + a) * ca
            if shift_atoms:
                base = taps[idx * atoms]
                assert taps[tap].dx == base.dx + a and taps[tap].dy == base.dy and taps[tap].c0 == base.c0
            for w in range(ca):
                meta.append((tap if tap < ntaps else -1, coff + w))
        mats = []
        for g in range(shifts):
            if shift_atoms:
                cols = [boxes[0][a:a + 64] for a in range(atoms)]
            else:
                cols = [b[g:g + 64] for b in boxes]
            mats.append(np.concatenate(cols, axis=1))
        return mats, meta

    for xi in range(NX):
        for yi in range(NY):
            acc = np.zeros((YS, M, N), dtype=np.float64)
            xmeta = ymeta = None
            for kt in range(total_kt):
                kx = kt % desc.kt_x
                r = kt // desc.kt_x
                ky = r % desc.kt_y
                n = r // desc.kt_y
                x0, y0 = kx * desc.tile_w, ky * desc.tile_h
                xt, xmeta = operand_tile(desc.x, xbuf, desc.x_taps, desc.x_taps_mode, desc.x_ntaps, desc.x_groups,
                                         desc.x_atoms, desc.x_ca, desc.x_shift_atoms, 1, xi, n, y0, x0)
                if SETS > 1:  # one shift-atom box per tap group yi*SETS + j
                    yt, ymeta = [], []
                    for j in range(SETS):
                        m_, meta_ = operand_tile(desc.y, ybuf, desc.y_taps, 1, desc.y_ntaps, desc.y_groups,
                                                 desc.y_atoms, desc.y_ca, 1, 1, yi * SETS + j, n, y0, x0)
                        yt.append(m_[0])
                        ymeta.append(meta_)
                else:
                    yt, meta_ = operand_tile(desc.y, ybuf, desc.y_taps, desc.y_taps_mode, desc.y_ntaps,
                                             desc.y_groups, desc.y_atoms, desc.y_ca, desc.y_shift_atoms, YSH, yi, n,
                                             y0, x0)
                    ymeta = [meta_] * YSH
                for g in range(YS):
                    acc[g] += xt[0].astype(np.float64).T @ yt[g].astype(np.float64)
            for g in range(YS):
                for m in range(M):
                    xtap, xch = xmeta[m]
                    for nn in range(N):
                        ytap, ych = ymeta[g][nn]
                        if xtap < 0 or ytap < 0:
                            continue
                        k, c = (xch, ych) if desc.x_is_dy else (ych, xch)
                        tap = desc.x_tap_rs[xtap] + desc.y_tap_rs[ytap] + (g if SETS == 1 else 0)
                        if tap < 0 or tap >= taps_total or k >= k_valid or c >= c_valid:
                            continue
                        dw[k * stride_k + c * stride_c + tap] = acc[g, m, nn]


def _run_wgrad_pair(desc, xbuf, ybuf, dw, stride_k, stride_c, k_valid, c_valid):
    """2-CTA wgrad (igemm_wgrad2_kernel): an item is a pair of taps; CTA r of the pair loads dy channels
    [128 r, +128) and input channels [128 r, +128) of both taps; D[256, 2 x 256]."""
    assert desc.x_is_dy and desc.x_ca == 64 and desc.y_ca == 64 and desc.x_atoms == 4 and desc.y_atoms == 4
    assert desc.y_sets == 2 and not desc.x_taps_mode and not desc.y_taps_mode and desc.stages <= 4
    assert int(desc.x.box[0]) == 64 and int(desc.x.box[1]) * int(desc.x.box[3]) == 64
    total_kt = desc.n_img * desc.kt_y * desc.kt_x
    for item in range((desc.y_ntaps + 1) // 2):
        acc = np.zeros((2, 256, 256), dtype=np.float64)
        for kt in range(total_kt):
            kx = kt % desc.kt_x
            r = kt // desc.kt_x
            ky = r % desc.kt_y
            n = r // desc.kt_y
            x0, y0 = kx * desc.tile_w, ky * desc.tile_h
            xt = np.concatenate([_tmap_gather(desc.x, xbuf, [rank * 128 + a * 64, x0, 0, y0, n]).reshape(64, 64)
                                 for rank in range(2) for a in range(2)], axis=1)
            for j in range(2):
                tap = item * 2 + j
                t = desc.y_taps[tap if tap < desc.y_ntaps else 0]
                yt = np.concatenate([_tmap_gather(desc.y, ybuf, [t.c0 + rank * 128 + a * 64, x0 + t.dx, t.plane,
                                                                 y0 + t.dy, n]).reshape(64, 64)
                                     for rank in range(2) for a in range(2)], axis=1)
                acc[j] += xt.astype(np.float64).T @ yt.astype(np.float64)
        for j in range(2):
            tap = item * 2 + j
            if tap >= desc.y_ntaps:
                continue
            tid = desc.y_tap_rs[tap]
            for k in range(min(256, k_valid)):
                for c in range(min(256, c_valid)):
                    dw[k * stride_k + c * stride_c + tid] = acc[j, k, c]
