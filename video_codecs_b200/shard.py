"""Multi-GPU partitioning of the path.  The path shares nothing across shards, so this is pure
host logic: which rank works on which frame pairs / which tile column, and how the per-rank MV fields are put back
in job order on the host.  No collective touches the data path; `gather_results` (used by tools and tests, not by the
benchmark's timed region) moves finished MV fields between host processes with whatever backend the process group has.

Names follow the reference's domain: frame pairs of a lookahead (BASELINE.json configs[4]), uniformly spaced tile
columns (configs[3]; TileUniformSpacing, TLibCommon/TComPicSym.cpp:217-229)."""
import numpy as np


def frame_pairs_of_rank(n_pairs, world, rank):
    """Contiguous block partition of pairs 0..n_pairs-1: lookahead pairs (t, t-1) are independent (originals only)."""
    base, extra = divmod(n_pairs, world)
    first = rank * base + min(rank, extra)
    return range(first, first + base + (1 if rank < extra else 0))


def tile_column_jobs(hm, pic_w, pic_h, n_columns, column, search_range=64, lambda_cost=0, pred=(0, 0), max_cu=64):
    """Canonical job list of one uniformly spaced tile column (HM tiles restrict MV prediction, not the search
    window, so a column's shard needs the whole reference plane but only its own jobs)."""
    x0, x1 = hm.tile_column_range(pic_w, n_columns, column, max_cu)
    ctus_y = (pic_h + max_cu - 1) // max_cu
    return hm.build_canonical_jobs_rect(pic_w, pic_h, (x0, x1), (0, ctus_y), search_range, lambda_cost, pred, max_cu)


def tile_column_crop(pic_w, ctu_x0, ctu_x1, search_range, max_cu=64, reach=4):
    """Picture columns [c0, c1) that the searches of the tile column's CTUs [ctu_x0, ctu_x1) can read: the column itself
    +- (search range + the 8-tap interpolation reach), clipped to the picture - SURVEY.md 8e's "column +- (R + 80) halo"
    without the part no window reaches.  Both ends are rounded out to whole CTUs: the kernels recognise the partitions of a
    CU by its alignment in plane coordinates, so a crop must start on a CTU boundary to keep every CU on the CU-fused path
    (any other start is still correct, but falls back to the per-PU / generic kernels).
    A rank uploads only this crop of both planes; beyond a crop edge that is not a picture edge the replicated margin holds
    wrong samples, which no candidate of the shard's jobs touches."""
    lo = ctu_x0 * max_cu - search_range - reach
    hi = min(pic_w, ctu_x1 * max_cu) + search_range + reach
    c0 = max(0, (lo // max_cu) * max_cu)
    c1 = min(pic_w, -(-hi // max_cu) * max_cu)
    return c0, c1


def shift_jobs(jobs, dx, dy=0):
    """The same jobs in the coordinates of a plane whose sample (0, 0) is picture sample (dx, dy) (windows and predictors are
    displacements and do not change)."""
    out = jobs.copy()
    out["pu_x"] -= dx
    out["pu_y"] -= dy
    return out


def job_sort_key(jobs):
    """Stable identity of a PU job inside one picture (position, size): used to merge shards back."""
    return np.lexsort((jobs["h"], jobs["w"], jobs["pu_x"], jobs["pu_y"]))


def merge_shards(job_shards, result_shards, all_jobs):
    """Puts per-shard results back in the order of `all_jobs` (every job of all_jobs must occur in exactly one shard)."""
    jobs = np.concatenate(job_shards)
    res = np.concatenate(result_shards)
    if len(jobs) != len(all_jobs):
        raise ValueError(f"shards hold {len(jobs)} jobs, the picture has {len(all_jobs)}")
    a, b = job_sort_key(jobs), job_sort_key(all_jobs)
    if not np.array_equal(jobs[a], all_jobs[b]):
        raise ValueError("shards do not partition the job list")
    out = np.empty_like(res)
    out[b] = res[a]
    return out


def gather_results(jobs_local, results_local, dst=0):
    """Collects (jobs, results) of every rank on `dst` through torch.distributed (any backend; host objects)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [None] * world if rank == dst else None
    dist.gather_object((jobs_local, results_local), box, dst=dst)
    if rank != dst:
        return None
    return [b[0] for b in box], [b[1] for b in box]
