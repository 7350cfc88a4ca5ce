"""Multi-GPU sharding of the hot path: one process per GPU, independent units, no exchange
step during compute; the only collective is the gather of outputs.

    configs 1/3/4  shard CHANNELS / STREAMS          -> split_units
    config 2       shard one stream into row-aligned frame ranges (each rank reads its own ntaps-decim halo
                   straight from the source)          -> chain_shard
    config 5       shard overlap-save BLOCKS; a rank's input slice starts at first_block*ngood and carries
                   its own nh-1 sample halo           -> fastfir_shard

Pure index arithmetic (no device code); covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass


def split_units(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of n_units for `rank`; the first n_units % world ranks get one extra."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class ChainShard:
    row_lo: int          # first output row of this rank
    row_hi: int
    in_start: int        # first input sample this rank reads
    in_len: int          # samples it reads (its frames plus the ntaps - decim tail)


def chain_shard(n_in: int, ntaps: int, decim: int, nfft: int, k_avg: int, rank: int, world: int) -> ChainShard:
    adv = nfft * decim
    tile_in = (nfft - 1) * decim + ntaps
    frames = 0 if n_in < tile_in else (n_in - tile_in) // adv + 1
    rows = frames // k_avg
    lo, hi = split_units(rows, rank, world)
    n_fr = (hi - lo) * k_avg
    in_len = 0 if n_fr == 0 else (n_fr - 1) * adv + tile_in
    return ChainShard(lo, hi, lo * k_avg * adv, in_len)


@dataclass
class FastFirShard:
    block_lo: int
    block_hi: int
    in_start: int
    in_len: int
    out_start: int
    out_len: int


def fastfir_shard(n_in: int, nh: int, nfft: int, rank: int, world: int) -> FastFirShard:
    """Full overlap-save blocks only (kff_nocopy, tools/kiss_fastfir.c:192-206); the flush block, if any,
    belongs to the last rank's caller."""
    ngood = nfft - nh + 1
    blocks = 0 if n_in < nfft else (n_in - nfft) // ngood + 1
    lo, hi = split_units(blocks, rank, world)
    nb = hi - lo
    in_len = 0 if nb == 0 else (nb - 1) * ngood + nfft
    return FastFirShard(lo, hi, lo * ngood, in_len, lo * ngood, nb * ngood)


def gather_rows(local, world: int):
    """All-gather equally shaped per-rank output blocks along dim 0 (NCCL over NVLink on GPUs, gloo on CPU).
    Ranks with fewer rows must pad to the common shape first (see pad_rows)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def pad_rows(local, rows: int):
    import torch
    if local.shape[0] == rows:
        return local
    pad = torch.zeros((rows - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    return torch.cat([local, pad], dim=0)
