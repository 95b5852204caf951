"""
Multi-GPU plumbing of the E-step: LD blocks are independent inside a sweep (row j only reaches columns of its own
block, /root/reference/viprs/model/vi/e_step.hpp:389-392,421), so whole LD blocks are sharded across ranks and the
only exchange per EM iteration is the small table of M-step / ELBO sums.

Everything here is host logic (numpy + torch.distributed); it runs unchanged over NCCL on GPUs and over gloo on
CPUs (the gloo world-2 / world-3 tests in tests/test_em_host.py).
"""
import numpy as np


def find_blocks(ld_left_bound, ld_indptr):
    """
    First row of every independent LD block, plus M at the end (same rule as csrc/ld.cu): a new block starts at row j
    when no earlier row's column run reaches j.  Accepts the upper-triangular and the symmetric layout.
    """
    lb = np.asarray(ld_left_bound, dtype=np.int64)
    ip = np.asarray(ld_indptr, dtype=np.int64)
    M = lb.shape[0]
    if M == 0:
        return np.zeros(1, dtype=np.int64)
    ends = lb + (ip[1:] - ip[:-1])                       # exclusive end of each row's run
    ends = np.maximum(ends, np.arange(M) + 1)            # an empty row still covers itself
    reach = np.maximum.accumulate(ends)                  # furthest column reached by rows 0..j
    starts = np.flatnonzero(np.concatenate([[True], reach[:-1] <= np.arange(1, M)]))
    return np.concatenate([starts, [M]]).astype(np.int64)


def block_costs(block_rows):
    """Sweep cost of a dense LD block ~ B^2 / 2 stored entries (+ a per-row term)."""
    b = np.diff(np.asarray(block_rows, dtype=np.int64)).astype(np.float64)
    return 0.5 * b * b + 256.0 * b


def partition_blocks(costs, world):
    """
    Contiguous runs of blocks per rank, balanced by cost: boundaries at the cost quantiles.
    Returns (world + 1,) block indices; rank r owns blocks [out[r], out[r+1]).
    """
    costs = np.asarray(costs, dtype=np.float64)
    nb = costs.shape[0]
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    total = cum[-1]
    out = np.zeros(world + 1, dtype=np.int64)
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target))
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, nb)] - target):
            k -= 1
        out[r] = min(max(k, out[r - 1]), nb)
    out[world] = nb
    return out


def shard_genome(chroms, world):
    """
    chroms: {chrom: dict(ld_left_bound, ld_indptr, ...)} in the order the model concatenates them.
    Returns per rank a dict {chrom: (row_start, row_end)} of contiguous whole-block row ranges (possibly empty),
    balancing the total sweep cost over ranks.
    """
    keys = list(chroms.keys())
    rows, costs, owner = [], [], []
    for c in keys:
        br = find_blocks(chroms[c]["ld_left_bound"], chroms[c]["ld_indptr"])
        rows.append(br)
        costs.append(block_costs(br))
        owner.append(np.full(len(br) - 1, len(owner)))
    allc = np.concatenate(costs) if costs else np.zeros(0)
    cut = partition_blocks(allc, world)
    first_block = np.concatenate([[0], np.cumsum([len(c) for c in costs])])
    out = []
    for r in range(world):
        b0, b1 = int(cut[r]), int(cut[r + 1])
        mine = {}
        for ci, c in enumerate(keys):
            lo = max(b0, int(first_block[ci])) - int(first_block[ci])
            hi = min(b1, int(first_block[ci + 1])) - int(first_block[ci])
            if hi > lo:
                mine[c] = (int(rows[ci][lo]), int(rows[ci][hi]))
            else:
                mine[c] = (0, 0)
        out.append(mine)
    return out


def slice_chromosome(ch, r0, r1):
    """Rows [r0, r1) of one chromosome's inputs as a self-contained LD matrix (whole blocks only)."""
    ip = np.asarray(ch["ld_indptr"])
    lb = np.asarray(ch["ld_left_bound"])
    out = dict(ch)
    out["ld_data"] = np.asarray(ch["ld_data"])[int(ip[r0]):int(ip[r1])]
    out["ld_indptr"] = (ip[r0:r1 + 1] - ip[r0]).astype(ip.dtype)
    out["ld_left_bound"] = (lb[r0:r1] - r0).astype(np.int32)
    for k in ("std_beta", "n_per_snp"):
        out[k] = np.asarray(ch[k])[r0:r1]
    return out


class SumsExchange:
    """
    The one collective of an EM iteration: a single SUM all-reduce over a packed float64 buffer holding the
    (nseg, ncol, NSUMS) table of every rank's shard.  The MAX_DIFF slot is a maximum, not a sum: each rank writes it
    into its own column of a (world, nseg, ncol) one-hot region, so one SUM collective carries it exactly.
    """

    def __init__(self, nseg, ncol, nsums, max_slot, rank=0, world=1, device="cpu", group=None):
        import torch
        self.torch = torch
        self.nseg, self.ncol, self.nsums, self.max_slot = nseg, ncol, nsums, max_slot
        self.rank, self.world, self.group = rank, world, group
        self.n_main = nseg * ncol * nsums
        self.n_max = world * nseg * ncol
        # send: [this rank's table | one-hot maxima] -- the rows of the other ranks stay zero for ever, so nothing is
        # cleared per iteration; recv: the copy the (in-place) collective works on, read back with ONE transfer
        self.send = torch.zeros(self.n_main + self.n_max, dtype=torch.float64, device=device)
        self.recv = torch.zeros_like(self.send)
        self.host = torch.zeros_like(self.send, device="cpu")
        if torch.device(device).type == "cuda":
            self.host = self.host.pin_memory()

    def table(self):
        """(nseg, ncol, nsums) view of the send buffer: let the sums kernel write here and all_reduce() copies nothing."""
        return self.send[:self.n_main].view(self.nseg, self.ncol, self.nsums)

    def onehot(self):
        """(world, nseg, ncol) view of the reduced per-rank maxima (valid after reduce_on_device)."""
        return self.recv[self.n_main:].view(self.world, self.nseg, self.ncol)

    def reduce_on_device(self, sums):
        """
        The collective alone, nothing read back: returns the (nseg, ncol, nsums) device table every rank's
        viprs_b200_em_update consumes (its MAX_DIFF slot is void for world > 1: use onehot()).  Stream-ordered, so it
        can be captured into a CUDA graph together with the kernels around it.
        """
        if self.world == 1:
            return sums
        import torch.distributed as dist
        main = self.table()
        if sums.data_ptr() != main.data_ptr():
            main.copy_(sums.view(self.nseg, self.ncol, self.nsums))
        self.send[self.n_main:].view(self.world, self.nseg, self.ncol)[self.rank].copy_(main[:, :, self.max_slot])
        self.recv.copy_(self.send)
        dist.all_reduce(self.recv, op=dist.ReduceOp.SUM, group=self.group)
        return self.recv[:self.n_main].view(self.nseg, self.ncol, self.nsums)

    def all_reduce(self, sums):
        """sums: (nseg, ncol, nsums) float64 tensor (device of the exchange) -> reduced numpy array."""
        if self.world == 1:
            if not sums.is_cuda:
                return sums.detach().numpy().reshape(self.nseg, self.ncol, self.nsums).copy()
            h = self.host[:self.n_main]
            h.copy_(sums.reshape(-1))                                  # blocking read-back into pinned memory
            return h.numpy().reshape(self.nseg, self.ncol, self.nsums).copy()
        import torch.distributed as dist
        main = self.table()
        if sums.data_ptr() != main.data_ptr():
            main.copy_(sums.view(self.nseg, self.ncol, self.nsums))
        self.send[self.n_main:].view(self.world, self.nseg, self.ncol)[self.rank].copy_(main[:, :, self.max_slot])
        self.recv.copy_(self.send)
        dist.all_reduce(self.recv, op=dist.ReduceOp.SUM, group=self.group)
        self.host.copy_(self.recv)                                   # synchronises (pageable or pinned destination)
        if self.recv.is_cuda:
            self.torch.cuda.current_stream().synchronize()
        h = self.host.numpy()
        out = h[:self.n_main].reshape(self.nseg, self.ncol, self.nsums).copy()
        # the summed MAX_DIFF slot of the main table is meaningless: take the maximum over the one-hot rows
        out[:, :, self.max_slot] = h[self.n_main:].reshape(self.world, self.nseg, self.ncol).max(axis=0)
        return out
