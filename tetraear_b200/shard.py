"""Carrier sharding over ranks (one process per GPU) and the one collective of the path.

Every (carrier, block) is independent in the reference (``SignalProcessor.process`` keeps no state
across calls, tetraear/signal/processor.py:221-273; one processor per capture device,
tetraear/ui/modern.py:1879), so carriers are block-partitioned over the ranks, each rank demodulates
only its own IQ, and the decoded dibit streams are exchanged once with an all-gather so that every rank
can run the host-side ``TetraDecoder`` on all carriers (BASELINE.json north_star).

torch.distributed is plumbing here (NCCL over NVLink on the GPU box, gloo in the CPU tests); the only
arithmetic in this module is the host-tensor form of the 2-bit transport packing used by those tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def partition(n_carriers: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first carrier, number of carriers) owned by `rank`: contiguous blocks, the first
    ``n_carriers % world_size`` ranks hold one carrier more."""
    if world_size <= 0 or not 0 <= rank < world_size or n_carriers < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(n_carriers, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def owner_of(carrier: int, n_carriers: int, world_size: int) -> int:
    """Rank that owns `carrier` under `partition`."""
    base, extra = divmod(n_carriers, world_size)
    split = extra * (base + 1)
    if carrier < split:
        return carrier // (base + 1)
    return extra + (carrier - split) // max(base, 1)


def gather_dibits(dibits: torch.Tensor, n_dibits: torch.Tensor, n_carriers: int,
                  out_dibits: Optional[torch.Tensor] = None, out_n: Optional[torch.Tensor] = None,
                  group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the fixed-stride dibit streams ``[c_local, cap]`` uint8 and their lengths ``[c_local]``
    int32 of every rank into ``[n_carriers, cap]`` / ``[n_carriers]`` on every rank (carrier order = rank
    order = global carrier index). With an even partition this is one ``all_gather_into_tensor`` per
    tensor (NCCL: a single ncclAllGather); a ragged partition pads every rank to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dibits, n_dibits
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    first, count = partition(n_carriers, world, rank)
    if dibits.shape[0] != count or n_dibits.shape[0] != count:
        raise ValueError(f"rank {rank} holds {dibits.shape[0]} carriers, partition says {count}")
    cap = dibits.shape[1]
    if out_dibits is None:
        out_dibits = torch.empty((n_carriers, cap), dtype=dibits.dtype, device=dibits.device)
    if out_n is None:
        out_n = torch.empty((n_carriers,), dtype=n_dibits.dtype, device=n_dibits.device)
    if n_carriers % world == 0:
        dist.all_gather_into_tensor(out_dibits, dibits.contiguous(), group=group)
        dist.all_gather_into_tensor(out_n, n_dibits.contiguous(), group=group)
        return out_dibits, out_n
    biggest = partition(n_carriers, world, 0)[1]
    pad_d = torch.zeros((biggest, cap), dtype=dibits.dtype, device=dibits.device)
    pad_n = torch.zeros((biggest,), dtype=n_dibits.dtype, device=n_dibits.device)
    pad_d[:count] = dibits
    pad_n[:count] = n_dibits
    all_d = torch.empty((world * biggest, cap), dtype=dibits.dtype, device=dibits.device)
    all_n = torch.empty((world * biggest,), dtype=n_dibits.dtype, device=n_dibits.device)
    dist.all_gather_into_tensor(all_d, pad_d, group=group)
    dist.all_gather_into_tensor(all_n, pad_n, group=group)
    for r in range(world):
        f, c = partition(n_carriers, world, r)
        out_dibits[f:f + c] = all_d[r * biggest: r * biggest + c]
        out_n[f:f + c] = all_n[r * biggest: r * biggest + c]
    return out_dibits, out_n


def _pack_cpu(d: torch.Tensor) -> torch.Tensor:
    """Reference packing for host tensors (the gloo tests): four dibits per byte, first dibit in the low bits."""
    q = d.reshape(-1, 4).to(torch.int32)
    return (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).to(torch.uint8)


def _unpack_cpu(p: torch.Tensor) -> torch.Tensor:
    q = p.reshape(-1, 1).to(torch.int32)
    return torch.cat([(q >> s) & 3 for s in (0, 2, 4, 6)], dim=1).to(torch.uint8).reshape(-1)


class PackedStreams:
    """The dibit streams and their lengths of one rank in ONE buffer, so that the exchange is a single collective, with
    the dibits packed four to a byte for the wire: ``[n_local * cap / 4]`` bytes followed by ``[n_local]`` int32 lengths
    (cap is rounded up to a multiple of 16). ``dibits`` / ``n_dibits`` are what the demodulator writes into;
    ``gather()`` packs, all-gathers the buffer and unpacks every rank's block. On the GPU the packing runs in the
    library's kernels (``packer`` = the ``SignalProcessor`` whose stream the demodulation runs on)."""

    def __init__(self, n_carriers: int, cap: int, device=None, group=None, packer=None, transport: str = "nccl"):
        """transport "nccl": pack kernel -> ``all_gather_into_tensor`` -> unpack kernel (gloo on CPU tensors);
        "p2p": the library's own exchange over NVLink peer memory (``tetra_allgather_dibits``: two kernels, no NCCL call) --
        the ranks' receive buffers are opened with CUDA IPC, the handles travel once through ``all_gather`` here."""
        self.group = group
        self.transport = transport
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if n_carriers % self.world:
            raise ValueError("PackedStreams needs an even partition; use gather_dibits for ragged ones")
        self.n_carriers, self.n_local = n_carriers, n_carriers // self.world
        self.cap = (cap + 15) & ~15
        self.packer = packer
        self.packed_bytes = self.n_local * self.cap // 4
        self.block = self.packed_bytes + 4 * self.n_local
        self.dibits = torch.zeros((self.n_local, self.cap), dtype=torch.uint8, device=device)
        self.local = torch.zeros(self.block, dtype=torch.uint8, device=device)
        self.n_dibits = self.local[self.packed_bytes:].view(torch.int32)
        if self.world > 1:
            self.all = torch.zeros(self.block * self.world, dtype=torch.uint8, device=device)
            self.out = torch.zeros((self.world, self.n_local, self.cap), dtype=torch.uint8, device=device)
        if self.dibits.is_cuda and packer is None and self.world > 1:
            raise ValueError("PackedStreams on a GPU needs the SignalProcessor as packer")
        if transport not in ("nccl", "p2p"):
            raise ValueError("transport must be 'nccl' or 'p2p'")
        if transport == "p2p" and self.world > 1:
            if not self.dibits.is_cuda:
                raise ValueError("the peer-memory exchange needs device tensors")
            self.out_n = torch.zeros((self.world, self.n_local), dtype=torch.int32, device=device)
            block = (self.block + 15) & ~15
            handle = packer.p2p_create(self.rank, self.world, block)
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            every = torch.zeros(self.world * len(handle), dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(every, mine, group=group)
            packer.p2p_connect(bytes(every.cpu().tolist()))
            dist.barrier(group=group)                      # nobody pushes before every rank has opened every buffer

    def gather(self):
        """Pack, one all-gather, unpack. Returns dibits ``[world, n_local, cap]`` (carrier c is
        ``[c // n_local, c % n_local]``) and lengths ``[world, n_local]``."""
        if self.world == 1:
            return self.dibits.unsqueeze(0), self.n_dibits.unsqueeze(0)
        if self.transport == "p2p":
            self.packer.set_stream(torch.cuda.current_stream(self.dibits.device).cuda_stream)
            self.packer.allgather_dibits_device(self.dibits.data_ptr(), self.n_local * self.cap, self.n_dibits.data_ptr(), self.n_local,
                                                self.out.data_ptr(), self.out_n.data_ptr())
            return self.out, self.out_n
        if self.dibits.is_cuda:
            # the library's pack / unpack kernels and the collective must be ordered on ONE stream: torch's current one
            self.packer.set_stream(torch.cuda.current_stream(self.dibits.device).cuda_stream)
            self.packer.pack_dibits_device(self.dibits.data_ptr(), self.n_local * self.cap, self.local.data_ptr())
        else:
            self.local[: self.packed_bytes] = _pack_cpu(self.dibits)
        dist.all_gather_into_tensor(self.all, self.local, group=self.group)
        blocks = self.all.view(self.world, self.block)
        if self.dibits.is_cuda:
            self.packer.unpack_dibits_device(self.all.data_ptr(), self.world, self.packed_bytes, self.block,
                                             self.out.data_ptr(), self.n_local * self.cap)
        else:
            for r in range(self.world):
                self.out[r] = _unpack_cpu(blocks[r, : self.packed_bytes]).view(self.n_local, self.cap)
        return self.out, blocks[:, self.packed_bytes:].view(torch.int32)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a per-rank scalar (the timing rule: a multi-GPU step takes as long as its slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
