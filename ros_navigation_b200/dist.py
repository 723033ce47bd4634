"""Multi-GPU plumbing for the batched mode: robots are independent, so they are block-partitioned over the ranks
(one process per GPU) with NO collective on the data path; once per cycle the 16-byte steering commands
(b200nav_command: speed, turnrate, picked angle, flags) of all robots are all-gathered so that every rank sees the
whole fleet (north_star: "a single NCCL-over-NVLink allgather of per-robot steering commands and statistics per
cycle, and only in batched mode").  The reference has no counterpart (one robot, one process; SURVEY section 2 rows
18-19).  torch.distributed is only plumbing here: NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

COMMAND_BYTES = 16


def partition(total, rank, world):
    """Contiguous block of robots owned by `rank`: [lo, hi).  The first total % world ranks get one more."""
    per, rem = divmod(total, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def owner_of(robot, total, world):
    """Rank that owns `robot` under partition()."""
    per, rem = divmod(total, world)
    edge = rem * (per + 1)
    return robot // (per + 1) if robot < edge else rem + (robot - edge) // max(per, 1)


class CommandExchange:
    """Per-cycle all-gather of the local robots' command records into a [total, 16] uint8 table."""

    def __init__(self, total, device, group=None, ctx=None, peer_push=False):
        """ctx: a capi.Context.  When given (CUDA, world > 1, equal shares) the gathers go through the library's own
        NCCL binding (b200nav_fleet_*): a handful of driver calls per cycle instead of a torch.distributed collective,
        which matters when the host, not the GPU, is the bottleneck of a sub-millisecond cycle."""
        self.total = total
        self.group = group
        self.fleet = None
        self.ctx = ctx
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = partition(total, self.rank, self.world)
        self.n_local = self.hi - self.lo
        self.max_local = -(-total // self.world)
        self.even = total % self.world == 0
        # double-buffered so that the gather of cycle k can overlap the kernels of cycle k+1
        self.locals = [torch.zeros(self.n_local, COMMAND_BYTES, dtype=torch.uint8, device=device) for _ in range(2)]
        self.tables = [torch.zeros(total, COMMAND_BYTES, dtype=torch.uint8, device=device) for _ in range(2)]
        self.local, self.table = self.locals[0], self.tables[0]
        self._pending = [None, None]
        self.push = False
        if ctx is not None and self.world > 1 and self.even and torch.device(device).type == "cuda":
            self._init_fleet(device)
            if peer_push:
                self._init_push(device)
        if not self.even:  # padded staging so that every rank contributes the same number of bytes
            self._send = torch.zeros(self.max_local, COMMAND_BYTES, dtype=torch.uint8, device=device)
            self._recv = torch.zeros(self.world * self.max_local, COMMAND_BYTES, dtype=torch.uint8, device=device)

    def _init_fleet(self, device):
        import ctypes as C
        from .capi import check, lib
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            check(lib().b200nav_fleet_unique_id(buf), None)
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(device)
        dist.broadcast(ident, src=0, group=self.group)
        raw = bytes(ident.cpu().tolist())
        h = C.c_void_p()
        check(lib().b200nav_fleet_create(self.ctx.h, raw, self.rank, self.world, C.byref(h)), self.ctx.h)
        self.fleet = h

    def _init_push(self, device):
        """Fused exchange: the VFH+ kernel stores the commands into every rank's table over NVLink peer mappings
        (b200nav_fleet_push_*); tables[slot] become views of the library's IPC-shared region."""
        import ctypes as C
        from .capi import check, lib
        buf = (C.c_uint8 * 64)()
        check(lib().b200nav_fleet_push_region(self.fleet, self.n_local, self.total, self.lo, buf), self.ctx.h)
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=device)
        allh = torch.zeros(64 * self.world, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        raw = bytes(allh.cpu().tolist())
        check(lib().b200nav_fleet_push_connect(self.fleet, raw), self.ctx.h)

        class _View:  # minimal __cuda_array_interface__ holder
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n, COMMAND_BYTES), "typestr": "|u1", "data": (ptr, False),
                                                 "version": 2, "strides": None}
        self.tables = [torch.as_tensor(_View(lib().b200nav_fleet_table(self.fleet, s), self.total), device=device)
                       for s in (0, 1)]
        self.table = self.tables[0]
        self.push = True

    def vfh_update_push(self, vfh, grid, layer, dev_inputs, slot):
        """The batched VFH+ update of this cycle with the exchange fused in; then wait(slot) before reading
        tables[slot]."""
        from .capi import check, lib, ptr
        self.wait(slot)
        check(lib().b200nav_vfh_update_batched_dev_push(vfh.h, grid.h, layer.encode(), ptr(dev_inputs), self.fleet, slot),
              self.ctx.h)
        self._pending[slot] = True

    def release(self, slot):
        """Peer push only: this rank has finished reading tables[slot] (reads enqueued on the context's stream so
        far).  Lets the other ranks start the slot's next cycle; without it the next vfh_update_push releases the slot
        itself, which keeps the ranks in lock step."""
        if self.push:
            from .capi import check, lib
            check(lib().b200nav_fleet_release(self.fleet, slot), self.ctx.h)

    def status(self):
        """Synchronise and raise if a peer-push wait timed out since the last call."""
        if self.fleet is not None:
            from .capi import check, lib
            check(lib().b200nav_fleet_status(self.fleet), self.ctx.h)

    def verify(self, slot):
        """Driver-visible proof of the exchange (call outside timed regions, after wait(slot)): on every rank,
        tables[slot] must be the concatenation of all ranks' rows of that cycle.  Each rank checksums the rows it
        produced (position-weighted, so swapped rows or ranks show) and every rank's block of the table it received;
        the per-block checksums of all ranks are gathered and compared with the producers' own.  Returns the number
        of (receiver, block) pairs that differ - 0 when the exchange delivered everything everywhere."""
        self.wait(slot)
        dev = self.tables[slot].device
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

        def checksum(rows, first_row):
            w = rows.contiguous().view(torch.int32).to(torch.int64).view(rows.shape[0], -1)
            idx = torch.arange(first_row, first_row + rows.shape[0], device=rows.device, dtype=torch.int64)[:, None]
            col = torch.arange(1, w.shape[1] + 1, device=rows.device, dtype=torch.int64)[None, :]
            return (w * (idx * 4 + col)).sum().reshape(1)

        own_rows = self.tables[slot][self.lo:self.hi] if self.push else self.locals[slot]
        produced = checksum(own_rows, self.lo)
        received = torch.cat([checksum(self.tables[slot][lo:hi], lo)
                              for lo, hi in (partition(self.total, r, self.world) for r in range(self.world))])
        if self.world == 1:
            return int((received != produced).sum())
        truth = torch.zeros(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(truth, produced, group=self.group)
        bad = (received != truth).sum().reshape(1)
        dist.all_reduce(bad, op=dist.ReduceOp.SUM, group=self.group)
        return int(bad)

    def close(self):
        if self.fleet is not None:
            from .capi import lib
            lib().b200nav_fleet_destroy(self.fleet)
            self.fleet = None

    def gather_async(self, slot):
        """Start the all-gather of `self.locals[slot]` into `self.tables[slot]` without blocking the caller's stream:
        the collective runs on the backend's own stream after the work already enqueued (NCCL), so it overlaps the
        next cycle's kernels.  Call wait(slot) before reading tables[slot] or rewriting locals[slot]."""
        self.wait(slot)
        if self.world == 1:
            self.tables[slot].copy_(self.locals[slot])
            return
        assert self.even, "gather_async needs total % world == 0"
        if self.fleet is not None:
            from .capi import check, lib
            check(lib().b200nav_fleet_gather_async(self.fleet, slot, self.locals[slot].data_ptr(),
                                                   self.tables[slot].data_ptr(), self.n_local * COMMAND_BYTES),
                  self.ctx.h)
            self._pending[slot] = True
            return
        self._pending[slot] = dist.all_gather_into_tensor(self.tables[slot].view(-1), self.locals[slot].view(-1),
                                                          group=self.group, async_op=True)

    def wait(self, slot=None):
        for s in ([slot] if slot is not None else [0, 1]):
            if self._pending[s] is not None:
                if self.fleet is not None:
                    from .capi import check, lib
                    check(lib().b200nav_fleet_wait(self.fleet, s), self.ctx.h)
                else:
                    self._pending[s].wait()
                self._pending[s] = None

    def gather(self):
        """All ranks call this once per cycle after their VFH+ kernel wrote `self.local`. Returns `self.table`."""
        if self.world == 1:
            self.table.copy_(self.local)
            return self.table
        if self.fleet is not None:
            self.gather_async(0)
            self.wait(0)
            return self.table
        if self.even:
            dist.all_gather_into_tensor(self.table.view(-1), self.local.view(-1), group=self.group)
            return self.table
        self._send[:self.n_local].copy_(self.local)
        dist.all_gather_into_tensor(self._recv.view(-1), self._send.view(-1), group=self.group)
        for r in range(self.world):
            lo, hi = partition(self.total, r, self.world)
            self.table[lo:hi].copy_(self._recv[r * self.max_local:r * self.max_local + (hi - lo)])
        return self.table
