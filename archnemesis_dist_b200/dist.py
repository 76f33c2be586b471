"""Multi-GPU sharding of the forward+Jacobian path: one process per GPU (torchrun), NCCL over
NVLink for the single exchange step.

The path shards three independent ways with no data-path collective until the end:
state-vector columns, geometries / paths, and wavenumbers (plus the (p,T) grid of the line-by-line generation).  Rank r takes the r-th contiguous chunk,
split exactly like the reference splits its joblib workers (archnemesis/ForwardModel_0.py:2322-2330:
``base = n // R`` with the first ``n % R`` chunks one longer), computes it against a full replica of
the k-table, and one ``all_gather`` assembles YN / KK on every rank for the (replicated, tiny)
optimal-estimation update.  There is nothing to fuse the collective with -- it is the last step --
so it is issued on the compute stream.
"""
import numpy as np
import torch
import torch.distributed as dist


def chunk_bounds(n, world):
    """[(lo, hi)] per rank; the reference's joblib split (ForwardModel_0.py:2322-2330)."""
    base, rem = divmod(n, world)
    return [(i * base + min(i, rem), (i + 1) * base + min(i + 1, rem)) for i in range(world)]


def my_chunk(n, rank=None, world=None):
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    return chunk_bounds(n, world)[rank]


def all_gather_rows(local, n_total, dim=0):
    """Assemble a tensor sharded along `dim` in contiguous, possibly ragged chunks (chunk_bounds) on
    every rank.  Equal chunks use one all_gather_into_tensor; ragged chunks are padded to the longest
    chunk first (the payload is NY*NX*8/R bytes per rank: latency-, not bandwidth-bound)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    bounds = chunk_bounds(n_total, world)
    longest = max(hi - lo for lo, hi in bounds)
    x = local.movedim(dim, 0).contiguous()
    if x.shape[0] != longest:
        pad = torch.zeros((longest - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    out = torch.empty((world * longest,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x)
    parts = [out[r * longest: r * longest + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0).movedim(0, dim).contiguous()


def jacobian_columns(hotpath, ev, M, to_tensor=None):
    """State-vector-column sharding: every rank evaluates the opacity and the layer-space Jacobian,
    projects only its own slice of columns (the projection is linear in xmap) and the KK columns are
    all-gathered.  Returns spec[NWAVE,NPATH], dspec_x[NWAVE,NPATH,NX] (full) as tensors."""
    NX = M.shape[2]
    lo, hi = my_chunk(NX)
    spec, dx_local, dtsurf = hotpath.forward_jacobian(ev, np.ascontiguousarray(M[:, :, lo:hi]))
    if to_tensor is not None:
        spec, dx_local, dtsurf = to_tensor(spec), to_tensor(dx_local), to_tensor(dtsurf)
    return spec, all_gather_rows(dx_local, NX, dim=2), dtsurf


def geometries(evaluate, n_geom):
    """Geometry sharding: `evaluate(i)` returns (spec_i[NWAVE], dspec_i[NWAVE,NX]) tensors for geometry
    i; rank r evaluates its chunk and YN[NGEOM,NWAVE], KK[NGEOM,NWAVE,NX] are all-gathered."""
    lo, hi = my_chunk(n_geom)
    specs, jacs = [], []
    for i in range(lo, hi):
        s, j = evaluate(i)
        specs.append(s)
        jacs.append(j)
    if specs:
        S, J = torch.stack(specs), torch.stack(jacs)
    else:
        raise ValueError("fewer geometries than ranks: give every rank at least one geometry")
    return all_gather_rows(S, n_geom, dim=0), all_gather_rows(J, n_geom, dim=0)


def pt_grid(absorption, pts, **kw):
    """(p,T)-grid sharding of the line-by-line cross-section generation (BASELINE config 3; the reference's
    calc_lbltable farms chunks of the grid out the same way, Spectroscopy_0.py:3124-3336): every state point is an
    independent slice of the launch, so rank r evaluates its contiguous chunk of `pts` with the full line list and
    the k[NPT, NWAVE] rows are all-gathered.  `absorption(pts_chunk)` returns the chunk's [n, NWAVE] tensor."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    lo, hi = my_chunk(len(pts))
    if hi <= lo:
        raise ValueError("fewer (p,T) points than ranks")
    return all_gather_rows(absorption(pts[lo:hi], **kw), len(pts), dim=0)


class WavenumberShard:
    """Wavenumber sharding: every stage of the path is independent per wavenumber, so rank r
    keeps only rows [lo, hi) of the k-table resident (1/R of the memory and of every kernel's work, nothing
    duplicated -- unlike geometry sharding, which repeats the gas opacity on every rank) and the ranks'
    [spectrum | Jacobian] row blocks are all-gathered.  The instrument line shape mixes neighbouring
    wavenumbers, so it is applied after the gather (or through forward_jacobian_conv on the gathered block).

    make_hotpath(K, PRESS, TEMP, DELG, WAVE) builds the rank's engine (engine.HotPath on the device)."""

    def __init__(self, make_hotpath, K, PRESS, TEMP, DELG, WAVE, rank=None, world=None):
        self.NWAVE = int(np.shape(K)[0])
        self.lo, self.hi = my_chunk(self.NWAVE, rank, world)
        if self.hi <= self.lo:
            raise ValueError("fewer wavenumbers than ranks")
        self.hotpath = make_hotpath(np.ascontiguousarray(K[self.lo:self.hi]), PRESS, TEMP, DELG,
                                    np.ascontiguousarray(np.asarray(WAVE)[self.lo:self.hi]))

    def slice_evaluation(self, ev):
        """The rank's rows of the per-wavenumber inputs of an Evaluation (continuum terms, emissivity, ...)."""
        import copy
        e = copy.copy(ev)
        for name in ("taucia", "taudust", "tauray", "dtaucon", "EMISSIVITY", "xfac", "SOLFLUX", "REFLECTANCE"):
            a = getattr(ev, name)
            if a is not None:
                setattr(e, name, np.ascontiguousarray(np.asarray(a)[self.lo:self.hi]))
        if getattr(ev, "continuum", None) is not None:
            # a continuum plan: the rank's rows of the resident CIA cross sections (cut once per table object) and of
            # the Rayleigh / aerosol spectra; the per-layer coefficients are shared
            from . import continuum as _cont
            tables, plan = ev.continuum
            cache = self.__dict__.setdefault("_cont_slices", {})
            mine = cache.get(id(tables))
            if mine is None or mine[0] is not tables:
                meta = dict(tables.meta, NWAVE=self.hi - self.lo)
                mine = (tables, _cont.ContinuumTables(tables.kw[:, :, self.lo:self.hi], tables.nplanes, meta))
                cache.clear()
                cache[id(tables)] = mine
            p = dict(plan)
            p["ur"] = np.ascontiguousarray(plan["ur"][:, self.lo:self.hi])
            p["ud"] = np.ascontiguousarray(plan["ud"][:, self.lo:self.hi])
            e.continuum = (mine[1], p)
        return e

    def forward_jacobian(self, ev, M, to_tensor=None):
        """spec[NWAVE,NPATH], dspec_x[NWAVE,NPATH,NX], dtsurf[NWAVE,NPATH] (full, on every rank)."""
        out = self.hotpath.forward_jacobian(self.slice_evaluation(ev), M)
        if to_tensor is not None:
            out = tuple(to_tensor(o) for o in out)
        return tuple(all_gather_rows(o, self.NWAVE, dim=0) for o in out)
