"""One particle filter whose particles are split across the GPUs of a torch.distributed group
(BASELINE.json configs[4]: R^4, 1e9 particles over 2/4/8 B200 with global systematic
resampling).  The reference has nothing comparable (single process, host memory).

Shard g owns the contiguous global particle range [g n_loc, (g+1) n_loc).  Per step:

  predict + weight   local (pfb_predict_loglik), no communication
  global max         all_reduce(MAX) of one double           -> every shard exponentiates against the same shift
  local CDF          sequential-exact scan of exp(logw - M)  (pfb_scan_cdf), shard total T_g
  shard offsets      all_gather of the G totals; O_{g+1} = fl(O_g + T_g) summed in rank order on every
                     rank (identical bits everywhere); global CDF v_i = fl(fl(O_g + c_i) / O_G)
  resampling         systematic comb u_j = (j + u0)/N over the GLOBAL CDF; output slot j lives on rank j // n_loc.
                     Pull form (default, pfb_resample_sharded_pull): every rank produces exactly its own n_loc
                     slots and reads the sources they come from wherever they live -- CDF shards and particle
                     buffers are torch symmetric memory, remote ones are read over NVLink inside the kernel, all
                     stores are local; the work per rank does not depend on where the weight sits.  Push form
                     (pull = False, pfb_resample_sharded_dev): the owner of a source stores its offspring into the
                     peers' buffers.  Without symmetric memory (gloo / CPU tests): host-made plan, pieces
                     [B_g, B_{g+1}) produced by shard g (pfb_resample_sharded) and one all_to_all_v.
  estimate           local sums + all_reduce(SUM)

This is exact global systematic resampling -- identical to a single filter holding all N particles, up to the
rounding of the shard-offset adds (oracle: pf_oracle.sharded_cdf / sharded_systematic_indices).  The
distributed logic below only talks to a small backend object, so it also runs on CPU tensors over gloo
(tests/test_sharded_gloo.py drives it with a numpy backend built from the oracle).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------------------------------
# host-side plan
# ----------------------------------------------------------------------------------------------------
def shard_offsets(totals):
    """O_0 = 0, O_{g+1} = fl(O_g + T_g), in rank order (float64, identical on every rank)."""
    offs = [0.0]
    for t in totals:
        offs.append(float(np.float64(offs[-1]) + np.float64(t)))
    return offs


def offspring_boundaries(offs, u0, n_out):
    """B_g = #{j in [0, n_out) : fl((j + u0)/n_out) < fl(O_g / O_G)} for g = 0..G; offspring
    [B_g, B_{g+1}) of the systematic comb are drawn from shard g."""
    total = np.float64(offs[-1])
    nn = np.float64(n_out)
    uu = np.float64(u0)
    B = []
    for g in range(len(offs)):
        E = np.float64(offs[g]) / total
        j = int(math.floor(float(E * nn - uu)))
        j = min(max(j, 0), n_out)
        while j < n_out and (np.float64(j) + uu) / nn < E:
            j += 1
        while j > 0 and (np.float64(j - 1) + uu) / nn >= E:
            j -= 1
        B.append(j)
    B[0] = 0
    B[-1] = n_out
    return B


def exchange_plan(B, rank, world, n_loc):
    """(send_counts, recv_counts) in particles: this rank generates offspring [B[rank], B[rank+1]) and
    output slot j lives on rank j // n_loc; it receives from source s the overlap of [B[s], B[s+1]) with
    its own slot range."""
    def overlap(a0, a1, b0, b1):
        return max(0, min(a1, b1) - max(a0, b0))

    send = [overlap(B[rank], B[rank + 1], r * n_loc, (r + 1) * n_loc) for r in range(world)]
    recv = [overlap(B[s], B[s + 1], rank * n_loc, (rank + 1) * n_loc) for s in range(world)]
    return send, recv


# ----------------------------------------------------------------------------------------------------
# the distributed resampling step (backend-agnostic)
# ----------------------------------------------------------------------------------------------------
def sharded_systematic_resample(backend, x, logw, local_max, u0, group=None, out=None, est_kind=None, peer_out=None):
    """x: (n_loc, D) local particles (already predicted), logw: (n_loc,), local_max: 1-element tensor.
    Returns the new local particles (n_loc, D).  `backend` provides
        scan(logw, shift) -> (cdf (n_loc,), total: 1-element tensor)
        offspring(cdf, offset, total, u0, j_begin, count, n_total, x, out=view, est_kind=None|"sum"|"trig")
            -> fills (count, D) rows; with est_kind returns the mean of those rows (or of their cos / sin)
    Collectives: all_reduce(MAX) of 1 double, all_gather of 1 double per rank, all_to_all_v of particles.
    peer_out(r) -> (n_loc, D) tensor aliasing rank r's `out` buffer (symmetric memory over NVLink): when given, the
    offspring kernels store the remote pieces straight into the peer's output slots -- no staging buffers, no
    all_to_all -- and the caller must run a collective on the same stream before any rank reads `out` (the
    estimate's all_reduce does)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_loc, D = x.shape
    n_total = n_loc * world
    gmax = local_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    cdf, total = backend.scan(logw, gmax)
    totals = [torch.empty_like(total) for _ in range(world)]
    dist.all_gather(totals, total, group=group)
    tot_host = [float(t.cpu()) for t in totals]  # the split sizes must be known on the host
    if not all(math.isfinite(t) for t in tot_host) or not sum(tot_host) > 0.0:
        raise AssertionError("The sum of all weights should be greater than 0.")
    offs = shard_offsets(tot_host)
    B = offspring_boundaries(offs, u0, n_total)
    send, recv = exchange_plan(B, rank, world, n_loc)
    if out is None:
        out = torch.empty((n_loc, D), dtype=x.dtype, device=x.device)
    lo_slot = rank * n_loc
    # sums over all offspring this rank PRODUCES (wherever they go): after an all_reduce they are the global sums
    est_sums = None
    if est_kind is not None:
        est_sums = torch.zeros(2 * D if est_kind == "trig" else D, dtype=torch.float64, device=x.device)

    def produce(j0, cnt, dst):
        piece = backend.offspring(cdf, offs[rank], offs[-1], u0, j0, cnt, n_total, x, out=dst, est_kind=est_kind)
        if est_kind is not None:
            est_sums.add_(piece * float(cnt))

    # the piece that stays on this rank is written straight into its output slots (no copy, no NCCL)
    if send[rank] > 0:
        j0 = max(B[rank], lo_slot)
        produce(j0, send[rank], out[j0 - lo_slot:j0 - lo_slot + send[rank]])
    # remote pieces: produced contiguously in destination-rank order, exchanged with one all_to_all_v
    send_remote = [0 if r == rank else send[r] for r in range(world)]
    recv_remote = [0 if r == rank else recv[r] for r in range(world)]
    if world > 1 and peer_out is not None:
        for r in range(world):
            if send_remote[r] > 0:
                j0 = max(B[rank], r * n_loc)
                produce(j0, send_remote[r], peer_out(r)[j0 - r * n_loc:j0 - r * n_loc + send_remote[r]])
        if est_kind is None:  # (with an estimate, the caller's all_reduce of the sums is the barrier)
            flag = torch.zeros(1, dtype=torch.float64, device=x.device)
            dist.all_reduce(flag, group=group)
    elif world > 1:
        sendbuf = torch.empty((sum(send_remote), D), dtype=x.dtype, device=x.device)
        pos = 0
        for r in range(world):
            if send_remote[r] > 0:
                j0 = max(B[rank], r * n_loc)
                produce(j0, send_remote[r], sendbuf[pos:pos + send_remote[r]])
                pos += send_remote[r]
        recvbuf = torch.empty((sum(recv_remote), D), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(recvbuf.view(-1), sendbuf.view(-1), [r * D for r in recv_remote], [s_ * D for s_ in send_remote],
                               group=group)
        pos = 0
        for s_ in range(world):
            if recv_remote[s_] > 0:
                j0 = max(B[s_], lo_slot)
                out[j0 - lo_slot:j0 - lo_slot + recv_remote[s_]].copy_(recvbuf[pos:pos + recv_remote[s_]])
                pos += recv_remote[s_]
    assert sum(recv) == n_loc
    return out, {"offsets": offs, "boundaries": B, "send": send, "recv": recv, "est_sums": est_sums}


def sharded_sum(local_sums, group=None):
    s = local_sums.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return s


# ----------------------------------------------------------------------------------------------------
# CUDA backend + the user-facing filter
# ----------------------------------------------------------------------------------------------------
class CudaShardBackend:
    def __init__(self, device):
        from ._engine import engine

        self.eng = engine(device)
        self._cdf = None

    def scan(self, logw, shift):
        from . import _lib as L

        eng = self.eng
        if self._cdf is None or self._cdf.shape != logw.shape:
            self._cdf = torch.empty_like(logw)
        eng.stats[L.STAT_MAX:L.STAT_MAX + 1].copy_(shift)  # the global shift replaces the shard's own max
        eng.scan(self._cdf, logw=logw, exact=True)
        return self._cdf, eng.stats[L.STAT_TOTAL:L.STAT_TOTAL + 1].clone()

    def offspring(self, cdf, offset, total, u0, j_begin, count, n_total, x, out=None, est_kind=None):
        """fills `out` with the offspring; returns the mean of the produced offspring (sum x, or per-dimension
        (cos, sin)) when est_kind is "sum" / "trig", else `out`"""
        from . import _lib as L

        D = x.shape[1]
        if out is None:
            out = torch.empty((count, D), dtype=x.dtype, device=x.device)
        if est_kind is None:
            if count > 0:
                self.eng.resample_sharded(cdf, offset, total, u0, j_begin, count, n_total, x, out, D)
            return out
        est = torch.zeros(2 * D if est_kind == "trig" else D, dtype=torch.float64, device=x.device)
        if count > 0:
            self.eng.resample_sharded(cdf, offset, total, u0, j_begin, count, n_total, x, out, D,
                                      est_kind=L.EST_TRIG if est_kind == "trig" else L.EST_SUM, est_out=est)
        return est


class ShardedParticleFilter:
    """A single filter over `n_particles_total` particles spread over the ranks of `group`.

        pf = ShardedParticleFilter("euclid", 10**9, 4)            # every rank, after init_process_group("nccl")
        pf.set_local_particles(x_local)                            # or pf.filter_state = prior (sampled per shard)
        pf.predict_nonlinear(LinearTransition(F), GaussianDistribution(0, Q))
        pf.update_identity(GaussianDistribution(0, R), z)          # global systematic resampling
        est = pf.get_point_estimate()                              # identical on every rank

    It wraps the single-GPU filter of the manifold for the local work (same kernels, same parameter blocks)."""

    def __init__(self, manifold, n_particles_total, dim, group=None, dtype=torch.float64, device=None):
        from . import filters as F

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if n_particles_total % self.world:
            raise ValueError("n_particles_total must be divisible by the number of shards")
        self.n_total = n_particles_total
        self.n_loc = n_particles_total // self.world
        cls = {"euclid": F.EuclideanParticleFilter, "torus": F.HypertoroidalParticleFilter,
               "sphere": F.HypersphericalParticleFilter}[manifold]
        self.manifold = manifold
        self.local = cls(self.n_loc, dim, dtype=dtype, device=device)
        self.local.lazy_predict = True
        self.backend = CudaShardBackend(self.local._filter_state.device)
        self.last_plan = None
        self._u0_rng = None
        self._sym = None        # [(tensor, handle)] x 2: the two particle buffers in symmetric memory (peer stores)
        self.peer_stores = True  # write remote offspring straight into the peer's buffer when symmetric memory works
        # device-side plan (pfb_resample_sharded_dev): one kernel per step, no host read of the shard totals; the
        # status of a step (global weight sum positive and finite) is checked when the NEXT step starts, or on check()
        self.device_plan = True
        # pull form (pfb_resample_sharded_pull): every rank produces its own n_loc slots and reads remote sources over
        # NVLink, so the work stays balanced when the weight sits in few shards; False = the push form (the owner of a
        # source stores its offspring into the peers' buffers)
        self.pull = True
        self._sym_cdf = None     # (tensor, handle): the CDF shard in symmetric memory (read by the peers in pull form)
        self._dp = None          # persistent small device / pinned tensors of the device-planned path
        self._pending_check = None

    def _symmetric_buffers(self, x):
        """Both particle buffers of the ping-pong in torch symmetric memory, so that every rank holds peer pointers
        to every other rank's output buffer.  Returns (x, peer_out) with x the current particles inside one of the
        two buffers, or (x, None) when symmetric memory is not available (CPU / gloo, one rank, unsupported box)."""
        loc = self.local
        if not self.peer_stores or self.world == 1 or x.device.type != "cuda":
            return x, None
        if self._sym is None:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                grp = self.group if self.group is not None else dist.group.WORLD
                bufs = []
                for _ in range(2):
                    t = symm_mem.empty(x.shape, dtype=x.dtype, device=x.device)
                    bufs.append((t, symm_mem.rendezvous(t, grp.group_name)))
                self._sym = bufs
            except Exception as exc:  # noqa: BLE001 -- fall back to the all_to_all exchange, once, loudly
                import warnings

                warnings.warn(f"symmetric memory unavailable ({exc!r}): sharded resampling exchanges through all_to_all")
                self.peer_stores = False
                return x, None
        (a, ha), (b, hb) = self._sym
        self._sym_handles = {a.data_ptr(): ha, b.data_ptr(): hb}
        if x.data_ptr() == a.data_ptr():
            cur, nxt, hn = a, b, hb
        elif x.data_ptr() == b.data_ptr():
            cur, nxt, hn = b, a, ha
        else:  # particles were (re)set from outside: move them into the symmetric pair
            a.copy_(x)
            cur, nxt, hn = a, b, hb
            st = loc._filter_state
            st._d = cur.reshape(st._d.shape)
        loc._alt = nxt
        shape, dtype = tuple(x.shape), x.dtype
        return cur, (lambda r: hn.get_buffer(r, shape, dtype))

    # -- state -------------------------------------------------------------------------------------------
    @property
    def filter_state(self):
        """the LOCAL shard as a Dirac distribution (weights 1/n_loc locally = 1/N globally after resampling)"""
        return self.local.filter_state

    def _rank_source(self):
        """the draw source of this shard: the global one when draws are injected, else its child number `rank` -- the
        in-kernel Philox draws are keyed by (seed, call, LOCAL particle index), so shards sharing a seed would draw
        identical noise"""
        from . import random as prandom

        src = prandom.source()
        return src.derive(self.rank) if isinstance(src, prandom.DeviceDraws) else src

    @filter_state.setter
    def filter_state(self, prior):
        from . import random as prandom

        self._est_sums = None
        with prandom.using(self._rank_source()):
            self.local.filter_state = prior  # every shard samples its own n_loc particles from its own stream

    def set_local_particles(self, x):
        self._est_sums = None
        self.local.filter_state.d = x  # (the setter clones a caller-owned device tensor: the filter mutates its buffers)

    def gather_particles(self):
        x = self.local.filter_state._records()
        parts = [torch.empty_like(x) for _ in range(self.world)]
        dist.all_gather(parts, x, group=self.group)
        return torch.cat(parts, dim=0)

    # -- predict / update ----------------------------------------------------------------------------------
    def predict_identity(self, noise_distribution):
        from .registry import IdentityTransition

        self.predict_nonlinear(IdentityTransition(), noise_distribution)

    def predict_nonlinear(self, f, noise_distribution=None, function_is_vectorized=True, shift_instead_of_add=None):
        from . import random as prandom

        self._est_sums = None
        with prandom.using(self._rank_source()):
            self.local.predict_nonlinear(f, noise_distribution, function_is_vectorized, shift_instead_of_add)

    def update_identity(self, meas_noise, measurement, shift_instead_of_add=True):
        from . import registry as R

        if not shift_instead_of_add:
            raise NotImplementedError()
        self.update_nonlinear_using_likelihood(meas_noise.set_mode(R._np(measurement)))

    def update_nonlinear_using_likelihood(self, likelihood, measurement=None, u0=None):
        import copy

        from . import _lib as L
        from . import random as prandom
        from . import registry as R

        loc = self.local
        dist_obj = R.as_distribution(likelihood)
        if measurement is not None:
            dist_obj = copy.deepcopy(dist_obj).set_mode(R._np(measurement))
        st = loc._filter_state
        x = loc._scratch()
        x, peer_out = self._symmetric_buffers(x)
        eng = loc._eng()
        lp, keep = R.lik_params(dist_obj, x.shape[1], images_to_device=R.device_put(x.device))
        if loc._pending is not None:
            p, noise = loc._pending
            loc._pending = None
            try:
                eng.predict_loglik(p, lp, x, x, noise, loc._logw)
            except NotImplementedError:
                eng.predict(p, x, x, noise)
                eng.loglik(lp, x, loc._logw)
        else:
            eng.loglik(lp, x, loc._logw)
        del keep
        local_max = eng.stats[L.STAT_MAX:L.STAT_MAX + 1].clone()
        if u0 is None:  # one comb offset for the whole filter, the same on every rank
            src = prandom.source()
            if src.fused:  # host generator seeded once from rank 0: no per-step broadcast / device sync
                if self._u0_rng is None:
                    t = torch.randint(0, 2**62, (1,), dtype=torch.int64, device=x.device)
                    dist.broadcast(t, src=0, group=self.group)
                    self._u0_rng = np.random.default_rng(int(t.cpu()))
                u0 = float(self._u0_rng.random())
            else:  # injected draws (parity tests): rank 0's uniform, broadcast
                t = src.uniforms((1,), x.device)
                dist.broadcast(t, src=0, group=self.group)
                u0 = float(t.cpu())
        self._last_u0 = u0
        if self.device_plan and x.device.type == "cuda" and (peer_out is not None or self.world == 1):
            self._update_device_plan(x, peer_out, local_max, u0)
            loc._alt = x  # ping-pong
            st._d = self._new.reshape(st._d.shape)
            st._uniform = True
            st._est_cache = None
            return
        new, plan = sharded_systematic_resample(self.backend, x, loc._logw, local_max, u0, self.group, out=loc._alt,
                                                est_kind="trig" if self.manifold == "torus" else "sum",
                                                peer_out=peer_out)
        self.last_plan = plan
        self._est_sums = sharded_sum(plan["est_sums"], self.group)  # global sums over all N offspring
        loc._alt = x  # ping-pong
        st._d = new.reshape(st._d.shape)
        st._uniform = True
        st._est_cache = None

    def _update_device_plan(self, x, peer_out, local_max, u0):
        """scan against the global shift, all-gather the shard totals, ONE scatter kernel that stores every offspring
        into the rank that owns its slot, all-reduce of the estimate sums (which is also the barrier before anyone
        reads the new particles).  The host reads nothing; a failed weight check surfaces at the next step / check()."""
        from . import _lib as L

        loc = self.local
        eng = loc._eng()
        n_loc, D = x.shape
        dev = x.device
        if self._dp is None:
            self._dp = {
                "totals": torch.zeros(self.world, dtype=torch.float64, device=dev),
                "est": torch.zeros(2 * L.PFB_MAXD, dtype=torch.float64, device=dev),
                "status": torch.zeros(1, dtype=torch.float64, device=dev),
                "host": torch.zeros(self.world + 1, dtype=torch.float64).pin_memory(),
                "peers": {},
            }
        dp = self._dp
        self.check()  # the previous step's status, copied to the host a step ago
        if self.world > 1:
            dist.all_reduce(local_max, op=dist.ReduceOp.MAX, group=self.group)
        pull = self.pull  # (one rank: the same kernel with a one-entry peer table)
        if pull and self.world > 1 and self._sym_cdf is None:
            import torch.distributed._symmetric_memory as symm_mem

            grp = self.group if self.group is not None else dist.group.WORLD
            t = symm_mem.empty((n_loc,), dtype=torch.float64, device=dev)
            self._sym_cdf = (t, symm_mem.rendezvous(t, grp.group_name))
        if pull and self.world > 1:
            self.backend._cdf = self._sym_cdf[0]  # the scan writes the shard's CDF where the peers can read it
        cdf, total = self.backend.scan(loc._logw, local_max)
        if self.world > 1:
            dist.all_gather_into_tensor(dp["totals"], total, group=self.group)  # (orders every peer's scan before the reads below)
        else:
            dp["totals"].copy_(total)
        out = loc._alt
        if pull:
            kx = ("pull", x.data_ptr())
            if kx not in dp["peers"]:
                if self.world == 1:
                    dp["peers"][kx] = ([cdf.data_ptr()], [x.data_ptr()])
                else:
                    hx, hc = self._sym_handles[x.data_ptr()], self._sym_cdf[1]
                    dp["peers"][kx] = (
                        [cdf.data_ptr() if r == self.rank else hc.get_buffer(r, (n_loc,), torch.float64).data_ptr() for r in range(self.world)],
                        [x.data_ptr() if r == self.rank else hx.get_buffer(r, tuple(x.shape), x.dtype).data_ptr() for r in range(self.world)])
            cdf_ptrs, x_ptrs = dp["peers"][kx]
            trig = self.manifold == "torus"
            dp["est"].zero_()
            timing = getattr(self, "time_scatter", False)
            if timing:
                self._scatter_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                self._scatter_ev[0].record()
            eng.resample_sharded_pull(cdf_ptrs, x_ptrs, dp["totals"], self.world, self.rank, u0, out, n_loc, D,
                                      L.EST_TRIG if trig else L.EST_SUM, dp["est"], dp["status"])
            if timing:
                self._scatter_ev[1].record()
            ne = 2 * D if trig else D
            sums = dp["est"][:ne].clone()
            if self.world > 1:  # (also the barrier: no rank overwrites a buffer while a peer may still be reading it)
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            self._finish_device_plan(sums, out, dp)
            return
        key = out.data_ptr()
        if key not in dp["peers"]:
            if self.world == 1:
                dp["peers"][key] = [out.data_ptr()]
            else:
                h = self._sym_handles[key]
                dp["peers"][key] = [out.data_ptr() if r == self.rank else h.get_buffer(r, tuple(x.shape), x.dtype).data_ptr()
                                    for r in range(self.world)]
        trig = self.manifold == "torus"
        dp["est"].zero_()
        timing = getattr(self, "time_scatter", False)
        if timing:
            self._scatter_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._scatter_ev[0].record()
        eng.resample_sharded_dev(cdf, dp["totals"], self.world, self.rank, u0, dp["peers"][key], x, D,
                                 L.EST_TRIG if trig else L.EST_SUM, dp["est"], dp["status"])
        if timing:
            self._scatter_ev[1].record()
        ne = 2 * D if trig else D
        sums = dp["est"][:ne].clone()
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)  # (also the barrier of the peer stores)
        self._finish_device_plan(sums, out, dp)

    def _finish_device_plan(self, sums, out, dp):
        self._est_sums = sums
        self._new = out
        self.last_plan = None
        self.last_totals = dp["totals"]
        # status + totals to pinned host memory, without waiting for them
        dp["host"][:1].copy_(dp["status"], non_blocking=True)
        dp["host"][1:].copy_(dp["totals"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(out.device))
        self._pending_check = ev

    def check(self):
        """raise the reference's AssertionError if the last device-planned update saw a non-positive / non-finite global
        weight sum (abstract_dirac_distribution.py:73-75); waits for that step's status copy only"""
        ev = self._pending_check
        if ev is None:
            return
        self._pending_check = None
        ev.synchronize()
        if float(self._dp["host"][0]) != 0.0:
            raise AssertionError("The sum of all weights should be greater than 0.")

    def scatter_ms(self):
        """device time of the last scatter kernel pair (time_scatter = True); synchronises"""
        a, b = self._scatter_ev
        b.synchronize()
        return a.elapsed_time(b)

    def offspring_traffic(self):
        """(offspring produced here, of which stored on other ranks) of the last device-planned update, from the totals
        it left on the device -- a host-side restatement of the plan for reporting, not on the step's path"""
        tot = [float(t) for t in self.last_totals.cpu()]
        offs = shard_offsets(tot)
        B = offspring_boundaries(offs, self._last_u0, self.n_total)
        send, recv = exchange_plan(B, self.rank, self.world, self.n_loc)
        if self.pull and self.world > 1:  # pull form: (offspring produced here = n_loc, of which from remote sources)
            return sum(recv), sum(recv) - recv[self.rank]
        return sum(send), sum(send) - send[self.rank]

    # -- estimate ------------------------------------------------------------------------------------------
    def get_point_estimate(self):
        from . import _lib as L

        loc = self.local
        if getattr(self, "_est_sums", None) is not None and loc._pending is None:
            s = self._est_sums  # reduced inside the resampling kernels: no extra pass over the particles
            D = loc._filter_state._stored_dim
            if self.manifold == "torus":
                t = s.reshape(D, 2)
                return torch.remainder(torch.atan2(t[:, 1], t[:, 0]), 2.0 * math.pi)
            m = s / self.n_total
            return m / torch.linalg.norm(m) if self.manifold == "sphere" else m
        st = loc.filter_state
        x = st._records()
        D = x.shape[1]
        if self.manifold == "torus":
            s = loc._eng().moments(L.MOM_TRIG, x, D)[: 2 * D].clone()
            s = sharded_sum(s, self.group).reshape(D, 2)
            return torch.remainder(torch.atan2(s[:, 1], s[:, 0]), 2.0 * math.pi)
        s = loc._eng().moments(L.MOM_MEAN, x, D)[:D].clone()
        s = sharded_sum(s, self.group) / self.world  # local sums carry weights 1/n_loc
        if self.manifold == "sphere":
            return s / torch.linalg.norm(s)
        return s
