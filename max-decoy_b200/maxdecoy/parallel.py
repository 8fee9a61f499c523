"""Multi-GPU layout of the identification hot path: spectra are sharded, the peptide index is replicated.

The reference processes spectra strictly serially (tasks/identification.rs:201) and every spectrum is
independent of the others, so the path shards by spectrum with no data-path collective.  The only exchange is
the gather of the fixed-width PSM tables at the end of a batch (one all_gather over NCCL/NVLink on GPUs,
gloo in the CPU tests).  Decoy RNG streams are keyed by the *global* spectrum id, so the gathered table does
not depend on the number of ranks.
"""
import numpy as np

from . import _abi

PSM_DTYPE = np.dtype(_abi.PSM_DTYPE)
PSM_BYTES = PSM_DTYPE.itemsize


def partition_spectra(precursor_mz, charge, world, block=32):
    """Index arrays, one per rank: spectra sorted by neutral precursor mass, dealt round-robin in blocks.

    Candidate counts grow ~10x with precursor mass, so dealing mass-sorted blocks gives every rank the same
    mass mix.  The union of the parts is a permutation of range(n); each part is ascending
    in mass, which also keeps neighbouring spectra on neighbouring index ranges."""
    mz = np.asarray(precursor_mz, dtype=np.float64)
    z = np.asarray(charge, dtype=np.float64)
    order = np.argsort(mz * z - 1.007276 * z, kind="stable")
    n = len(order)
    nblk = (n + block - 1) // block
    parts = [[] for _ in range(world)]
    for b in range(nblk):
        parts[b % world].append(order[b * block:(b + 1) * block])
    return [np.concatenate(p) if p else np.zeros(0, dtype=np.int64) for p in parts]


def shard(spectra, rank, world, block=32):
    """The rank's shard of `spectra`, carrying global spectrum ids."""
    part = partition_spectra(spectra.precursor_mz, spectra.charge, world, block)[rank]
    sub = spectra.subset(part)
    if spectra.spectrum_id is not None:
        sub.spectrum_id = np.ascontiguousarray(spectra.spectrum_id[part], dtype=np.uint32)
    else:
        sub.spectrum_id = np.ascontiguousarray(part, dtype=np.uint32)
    return sub, part


def padded_rows(n_total, world, block=32):
    """Rows every rank contributes to the all_gather (the largest shard); shorter shards pad with rank-0 rows."""
    nblk = (n_total + block - 1) // block
    most = (nblk + world - 1) // world
    return most * block


def gather_psms(psms, top_k, n_total, group=None, device=None, block=32):
    """all_gather of per-rank PSM tables ([n_local, top_k] structured arrays) -> [n_total, top_k] table ordered
    by spectrum id.  Uses torch.distributed (NCCL for CUDA tensors, gloo on CPU); with no process group it
    is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return order_by_spectrum(psms.reshape(-1, top_k), n_total)
    world = dist.get_world_size(group)
    rows = padded_rows(n_total, world, block)
    buf = np.zeros((rows, top_k), dtype=PSM_DTYPE)
    buf["spectrum_id"] = 0xFFFFFFFF          # padding rows
    buf[:len(psms)] = psms.reshape(-1, top_k)
    send = torch.from_numpy(buf.view(np.uint8).reshape(-1))
    if device is not None:
        send = send.to(device)
    recv = torch.empty(send.numel() * world, dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    allrows = recv.cpu().numpy().view(PSM_DTYPE).reshape(-1, top_k)
    keep = allrows["spectrum_id"][:, 0] != 0xFFFFFFFF
    return order_by_spectrum(allrows[keep], n_total)


def pad_rows(psms, top_k, rows):
    """[n_local, top_k] table -> [rows, top_k] with padding rows (spectrum_id 0xFFFFFFFF, rank 0) behind it."""
    buf = np.zeros((rows, top_k), dtype=PSM_DTYPE)
    buf["spectrum_id"] = 0xFFFFFFFF
    buf[:len(psms)] = psms.reshape(-1, top_k)
    return buf


def drop_padding(allrows, top_k, n_total):
    allrows = allrows.reshape(-1, top_k)
    keep = allrows["spectrum_id"][:, 0] != 0xFFFFFFFF
    return order_by_spectrum(allrows[keep], n_total)


def identify_sharded_comm(engine, spectra, params, rank, world, block=32):
    """The same over the library's own collective (md_comm_init / md_gather_psms of include/maxdecoy.h: NCCL inside the
    CUDA library, a copy in a single-rank run): what a host without torch.distributed (the Rust / C++ host) calls."""
    sub, _ = shard(spectra, rank, world, block)
    psms, stats = engine.identify(sub, params)
    rows = padded_rows(len(spectra), world, block)
    allrows = engine.gather_psms(pad_rows(psms, params.top_k, rows))
    return drop_padding(allrows, params.top_k, len(spectra)), stats


def order_by_spectrum(rows, n_total):
    """Sort gathered [n, top_k] rows by global spectrum id (stable)."""
    order = np.argsort(rows["spectrum_id"][:, 0], kind="stable")
    out = rows[order]
    assert len(out) == n_total, "gathered %d spectra, expected %d" % (len(out), n_total)
    return out


def identify_sharded(engine, spectra, params, group=None, device=None, block=32):
    """identification_task over this rank's shard + PSM gather.  Every rank returns the full table."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    sub, _ = shard(spectra, rank, world, block)
    psms, stats = engine.identify(sub, params)
    return gather_psms(psms, params.top_k, len(spectra), group, device, block), stats
