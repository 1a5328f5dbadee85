"""Host-side sharding helpers for multi-GPU runs (one process per GPU).

Events are independent (per-event Philox streams), so rank r of N simply owns its own event-id range; the only
exchange is the all-reduce of the binned-flux block whose trailing counters carry f_inx (MainRunner.jl:749).
"""
import numpy as np

N_TRAILING = 8  # [f_inx, n_rows, n_events, n_acc, n_rej, n_rhs, n_cond, n_rays] (include/art.h)


def event_range(rank: int, world: int, step: int, events_per_rank: int):
    """First event id and count owned by `rank` in bench/production step `step` (ids start at 1 like the
    reference's photon_trajs counter); ranges of different (step, rank) never overlap."""
    if not (0 <= rank < world) or events_per_rank < 0 or step < 0:
        raise ValueError("bad shard request")
    return 1 + (step * world + rank) * events_per_rank, events_per_rank


def split_events(n_total: int, world: int):
    """Contiguous split of events 1..n_total over `world` ranks (strong scaling): [(first, count), ...]."""
    base, rem = divmod(n_total, world)
    out, first = [], 1
    for r in range(world):
        c = base + (1 if r < rem else 0)
        out.append((first, c))
        first += c
    return out


def binned_flux(rows: np.ndarray, f_inx: int, n_costheta: int, n_phi: int, extra_counters=()):
    """numpy restatement of k_hist: bins [species][cosθf][ϕf][Σ w·sln, Σ (w·sln)², Σ w, count] + trailing counters."""
    H = np.zeros(2 * n_costheta * n_phi * 4 + N_TRAILING)
    b = H[:-N_TRAILING].reshape(2, n_costheta, n_phi, 4)
    if len(rows):
        sp = rows[:, 1].astype(int)
        ic = np.clip(np.floor((np.cos(rows[:, 2]) + 1.0) * 0.5 * n_costheta).astype(int), 0, n_costheta - 1)
        ip = np.clip(np.floor((rows[:, 3] + np.pi) / (2 * np.pi) * n_phi).astype(int), 0, n_phi - 1)
        w = rows[:, 8] * rows[:, 7]
        np.add.at(b[..., 0], (sp, ic, ip), w)
        np.add.at(b[..., 1], (sp, ic, ip), w * w)
        np.add.at(b[..., 2], (sp, ic, ip), rows[:, 8])
        np.add.at(b[..., 3], (sp, ic, ip), 1.0)
    H[-N_TRAILING] = f_inx
    H[-N_TRAILING + 1] = len(rows)
    for k, v in enumerate(extra_counters):
        H[-N_TRAILING + 2 + k] = v
    return H


def normalised_flux(H: np.ndarray, n_costheta: int, n_phi: int):
    """After the all-reduce: divide the first two moments by f_inx (f_inx²) — MainRunner.jl:749 applied to bins."""
    f = H[-N_TRAILING]
    b = H[:-N_TRAILING].reshape(2, n_costheta, n_phi, 4).copy()
    if f > 0:
        b[..., 0] /= f
        b[..., 1] /= f * f
    return b
