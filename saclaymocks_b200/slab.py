"""Host-side description of the slab decomposition: who owns what, and the buffer layouts on
both sides of the all-to-all.  chunk.py uses the ownership helpers; the layout helpers are the numpy statement of
the addressing the strided FFT kernels use (smk_capi.cu: PassAddr of smk_fft_r2c_local / smk_synth_c2r_finish) and
are exercised with a real collective (gloo) in tests/test_dist_cpu.py."""
import numpy as np


def x_planes(rank, nranks, NX):
    """Global x-planes [lo, hi) owned by `rank` (make_spectra.py:220-221 without the halo)."""
    return rank * NX // nranks, (rank + 1) * NX // nranks


def halo(rank, nranks, dmax):
    """Halo planes below / above the owned slab: global edges get none (make_spectra.py:220-221)."""
    return (dmax if rank > 0 else 0), (dmax if rank < nranks - 1 else 0)


def x_bounds(rank, nranks, LX):
    """Ownership interval xmin < X <= xmax of the slab (make_spectra.py:279-280, 443-448)."""
    return LX * rank / nranks - LX / 2, LX * (rank + 1) / nranks - LX / 2


def touching(xyzr, r_first, r_last, xmin, xmax):
    """Quasars whose sightline (X grows monotonically with R) can own a pixel of the slab."""
    xa = xyzr[:, 0] * r_first / xyzr[:, 3]
    xb = xyzr[:, 0] * r_last / xyzr[:, 3]
    return (np.minimum(xa, xb) <= xmax) & (np.maximum(xa, xb) > xmin)


def home_rank(xyzr, nranks, LX):
    """Rank whose slab contains the quasar itself (names the output file, like the reference's HDU index)."""
    return np.clip(np.floor((xyzr[:, 0] + LX / 2) / (LX / nranks)).astype(int), 0, nranks - 1)


# ---- all-to-all layouts (P = padded row length of boxk)
def forward_pack(local_xyk, nranks):
    """[nxl][NY][P] after the local z and y passes -> send buffer [dest][nxl][nyl][P] (dest = y // nyl)."""
    nxl, NY, P = local_xyk.shape
    nyl = NY // nranks
    return np.ascontiguousarray(local_xyk.reshape(nxl, nranks, nyl, P).transpose(1, 0, 2, 3))


def forward_unpack(recv, nranks):
    """Received [src][nxl][nyl][P] IS boxk's y-slab layout [NX][nyl][P] (x = src*nxl + xl): a reshape."""
    src, nxl, nyl, P = recv.shape
    return recv.reshape(src * nxl, nyl, P)


def inverse_pack(boxk_yslab, nranks):
    """[NX][nyl][P] after the x pass: destination chunks (x ranges) are already contiguous."""
    NX, nyl, P = boxk_yslab.shape
    return boxk_yslab.reshape(nranks, NX // nranks, nyl, P)


def inverse_unpack(recv):
    """Received [src][nxl][nyl][P] -> the x-slab [nxl][NY][P] the y pass reads (y = src*nyl + yl), done in the
    kernel by two-level addressing."""
    src, nxl, nyl, P = recv.shape
    return np.ascontiguousarray(recv.transpose(1, 0, 2, 3)).reshape(nxl, src * nyl, P)


# ---- gather of the sharded spectra rows to the home rank (the reference does this through the
# file system: make_spectra.py writes one piece per slice, merge_spectra.py:385-406 stitches the pieces of a quasar)
def row_gather_plan(xyzr, r_first, r_last, nranks, LX, rank):
    """Who sends which rows to whom.  xyzr: the FULL catalogue [n, 4].  A rank holds a row for every quasar whose
    sightline touches its slab (`touching`), in ascending catalogue order; the home rank of a quasar is the slab that
    contains the quasar itself (`home_rank`).  Returns a dict with, for `rank`:
      send_rows  indices into the local rows, ordered by destination then catalogue index,
      send_split / recv_split  rows per destination / per source,
      recv_qso   catalogue index of every received row (ordered by source then catalogue index),
      home_qso   catalogue indices this rank is home of (ascending)."""
    home = home_rank(xyzr, nranks, LX)
    touch = []
    for r in range(nranks):
        xmin, xmax = x_bounds(r, nranks, LX)
        touch.append(touching(xyzr, r_first, r_last, xmin, xmax))
    mine = np.where(touch[rank])[0]                       # catalogue indices of the local rows, ascending
    local_of = {int(q): i for i, q in enumerate(mine)}
    send_rows, send_split = [], []
    for d in range(nranks):
        qs = mine[home[mine] == d]
        send_rows += [local_of[int(q)] for q in qs]
        send_split.append(len(qs))
    recv_qso, recv_split = [], []
    for s in range(nranks):
        qs = np.where(touch[s] & (home == rank))[0]
        recv_qso += list(qs)
        recv_split.append(len(qs))
    return dict(send_rows=np.asarray(send_rows, dtype=np.int64), send_split=send_split, recv_split=recv_split,
                recv_qso=np.asarray(recv_qso, dtype=np.int64), home_qso=np.where(home == rank)[0])


def exchange_rows(rows, plan, group=None):
    """rows: torch tensor [n_local, width] (NaN where the slab does not own the pixel).  One all-to-all moves every
    row to its quasar's home rank; the pieces are merged by position (a pixel is owned by exactly one slab).  Returns
    [len(plan['home_qso']), width]; pixels nobody owns (outside the box) stay NaN.  Works for CPU (gloo) and CUDA (NCCL)
    tensors."""
    import torch
    import torch.distributed as dist
    width = rows.shape[1]
    idx = torch.as_tensor(plan["send_rows"], device=rows.device)
    send = rows.index_select(0, idx).contiguous() if len(idx) else rows.new_empty((0, width))
    recv = rows.new_empty((int(sum(plan["recv_split"])), width))
    dist.all_to_all_single(recv, send, output_split_sizes=list(plan["recv_split"]),
                           input_split_sizes=list(plan["send_split"]), group=group)
    home = plan["home_qso"]
    out = rows.new_full((len(home), width), float("nan"))
    if len(home) == 0 or recv.shape[0] == 0:
        return out
    slot = torch.as_tensor(np.searchsorted(home, plan["recv_qso"]), device=rows.device)
    o = 0
    for n in plan["recv_split"]:                          # one source at a time: rows of a source are distinct quasars
        if n:
            piece, where = recv[o:o + n], slot[o:o + n]
            cur = out.index_select(0, where)
            out.index_copy_(0, where, torch.where(torch.isnan(piece), cur, piece))
        o += n
    return out
