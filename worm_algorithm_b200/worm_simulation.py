"""`WormSimulation` -- the reference's interface (worm_simulation.py:8-188) on a B200.

Same constructor arguments, same result attributes (`_T`, `_beta`, `_Z`, `_E`, `_Nb`,
`_sim_steps`, dicts keyed `str(T).rstrip('0')` exactly like worm_simulation.py:119-126).  Where
the reference compiles a C++ file and `Popen`s it once per temperature, this class makes ONE
batched call into libworm_b200.so for all temperatures x replicas and gets tensors back.

Two modes:
  rng="mt19937"  deterministic: the reference's chain on the reference's MT19937 stream, bit-exact
                 (`seeds` are the doubles the reference writes into input.txt; if omitted they are
                 drawn like worm_simulation.py:46).  The values stored in the result dicts go
                 through the same "%.12f" text rounding output.txt imposes (worm_ising_2d.cpp:190).
  rng="philox"   production: counter-based RNG, fixed-length chains, warm-up excluded from the
                 averages; `_E` = -T<Nb>/N, `_Z` = closed fraction, `_Nb` = <Nb>/N.

`rng="philox"` is the production mode (1e11-1e12 worm steps/s on a B200); the default `rng="mt19937"`
is kept because it is what a literal drop-in call must reproduce, but it is the parity vehicle (one warp
per chain, ~5e6 steps/s per chain).

`data_dir` (default None = nothing written) reproduces the reference's `../data` tree
 in BOTH modes -- observables rows, bonds_{T}.txt, num_bonds, output.txt, bond map -- so its
analysis classes (`SpecificHeat`, `Bonds(run=False)`, `CountBonds`) read the results unchanged.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import compat_io, engine
from ._lib import ACC_ITERS, ACC_N, ACC_N2, ACC_NB, ACC_NB2, ACC_STATUS, ACC_Z, ALGO_BINARY, ALGO_TAYLOR


def temperature_key(T) -> str:
    """worm_simulation.py:121: `str(T).rstrip('0')`."""
    return str(float(T)).rstrip('0')


class WormSimulation(object):
    def __init__(self, L=32, run=True, num_steps=1E7, verbose=True,
                 T_start=1., T_end=3.5, T_step=0.1, T_arr=None, bond_flag=1,
                 seeds=None, n_replicas=1, rng="mt19937", algo="taylor", therm_frac=0.1,
                 write_steps=50, max_dumps=None, data_dir=None, device=None,
                 lanes_per_chain=0, rank=0, world_size=1, hist=False, gather_chains=False):
        self._L = int(L)
        self._num_steps = int(num_steps)
        self._verbose = verbose
        self._bond_flag = int(bond_flag)
        if T_arr is None:
            self._T_range = np.arange(T_start, T_end, T_step)        # worm_simulation.py:22
        else:
            self._T_range = np.asarray(T_arr, dtype=np.float64)
        self._run = run
        if rng not in ("mt19937", "philox"):
            raise ValueError("rng must be 'mt19937' or 'philox'")
        if algo not in ("taylor", "binary"):
            raise ValueError("algo must be 'taylor' or 'binary'")
        if rng == "mt19937" and algo != "taylor":
            raise ValueError("the deterministic mode is the reference's Taylor-expansion chain")
        self._rng = rng
        self._algo = ALGO_TAYLOR if algo == "taylor" else ALGO_BINARY
        self._n_replicas = int(n_replicas)
        self._therm_frac = float(therm_frac)
        self._write_steps = int(write_steps)
        self._max_dumps = max_dumps
        self._device = device
        self._lanes = int(lanes_per_chain)
        self._rank, self._world = int(rank), int(world_size)
        self._seeds = seeds
        self._hist = bool(hist)
        self._gather = bool(gather_chains)
        if rng == "mt19937" and self._write_steps != 50:
            raise ValueError("the deterministic mode dumps every 50 steps like worm_ising_2d.cpp:114; write_steps is a "
                             "production-mode (rng='philox') option")
        self._tree = compat_io.DataTree(data_dir, self._L) if data_dir is not None else None
        self.results = None
        if self._run:
            self.remove_old_data()
            self.run()

    # -- reference methods that have no work left to do -----------------------------------------
    def make(self):
        """The reference recompiles its C++ here (worm_simulation.py:69-89); the CUDA library is
        prebuilt, so this only checks that it loads."""
        engine.lib()

    def prepare_input(self, T, num_steps):
        """worm_simulation.py:44-57 draws `seed = T + randint(1e4) * rand()`; kept for callers that
        want the reference's (unseeded) seed law."""
        return T + np.random.randint(1E4) * np.random.rand()

    def remove_old_data(self):
        if self._tree is not None:
            self._tree.remove_old_bonds()

    def clean(self):
        pass

    # -------------------------------------------------------------------------------------------
    def _chain_table(self):
        """(T per chain, replica index per chain, temperature index per chain); temperatures fastest,
        so a contiguous shard of the chain range holds every temperature."""
        nT, R = len(self._T_range), self._n_replicas
        t_idx = np.tile(np.arange(nT), R)          # replica-major: chain c = replica * nT + t
        r_idx = np.repeat(np.arange(R), nT)
        return np.asarray(self._T_range, dtype=np.float64)[t_idx], r_idx, t_idx

    def run(self):
        self._T = []
        self._beta, self._Z, self._E, self._Nb, self._sim_steps = {}, {}, {}, {}, {}
        if len(self._T_range) == 0:
            return
        if self._rng == "mt19937":
            self._run_mt()
        else:
            self._run_philox()
        if self._tree is not None:
            self._tree.write_headers()

    # -- deterministic mode ---------------------------------------------------------------------
    def _run_mt(self):
        L = self._L
        Tc, r_idx, t_idx = self._chain_table()
        # the executable reads T from "%.12f" text (worm_simulation.py:53, worm_ising_2d.cpp:44)
        Tc = np.array([float("%.12f" % t) for t in Tc])
        n = Tc.size
        if self._seeds is None:
            seeds = np.array([self.prepare_input(t, self._num_steps) for t in Tc])
        else:
            seeds = np.asarray(self._seeds, dtype=np.float64).reshape(-1)
            if seeds.size == len(self._T_range) and self._n_replicas > 1:
                seeds = seeds[t_idx] + 10007.0 * r_idx
            if seeds.size != n:
                raise ValueError("seeds: one per temperature (x replica)")
        # "%.32f" text round trip of the seed is exact for doubles of this size (worm_simulation.py:53)
        want_dumps = bool(self._bond_flag)
        max_events = self._max_dumps
        if max_events is None:
            # The executable keeps EVERY event (a closed iteration at a multiple of 50 after the warm-up,
            # worm_ising_2d.cpp:113-125): size the buffers from a counting pass (max_events = 0 stores nothing
            # and returns n_events; the chains are deterministic, the second pass repeats them).  Only asked
            # for when the events are used -- for the files, or for the dumps.
            if self._tree is not None or want_dumps:
                cnt = engine.run_mt(L, Tc, seeds, self._num_steps, bond_flag=0, max_events=0, device=self._device)
                max_events = int(cnt.n_events.max().item()) if n else 0
                per_event = 24 + (2 * L * L if want_dumps else 0)
                if n * max_events * per_event > (8 << 30):
                    raise MemoryError("deterministic mode: {} chains x {} events x {} B exceed 8 GiB; pass max_dumps=... "
                                      "to keep only the first events of every chain".format(n, max_events, per_event))
            else:
                max_events = 0
        res = engine.run_mt(L, Tc, seeds, self._num_steps, bond_flag=self._bond_flag,
                            max_events=max_events, device=self._device)
        out = res.out.cpu().numpy()
        n_ev = res.n_events.cpu().numpy()
        self.results = dict(T=Tc, seeds=seeds, replica=r_idx, out=res.out, bonds=res.bonds,
                            n_events=res.n_events, events=res.events, dumps=res.dumps, mode="mt19937")
        if (out[:, 4] != 0).any():
            raise RuntimeError("bond occupation exceeded 254 (uint8 storage)")
        self.results["n_events_dropped"] = int(np.maximum(n_ev - max_events, 0).sum())
        if self.results["n_events_dropped"]:
            warnings.warn("deterministic mode: {} events beyond max_dumps={} per chain were not stored; the files hold "
                          "the first {} of each chain (the executable writes all)".format(
                              self.results["n_events_dropped"], max_events, max_events))
        events = res.events.cpu().numpy() if (self._tree is not None) else None
        dumps = res.dumps.cpu().numpy() if (self._tree is not None and want_dumps and res.dumps is not None) else None
        final = res.bonds.cpu().numpy() if (self._tree is not None and want_dumps) else None
        N = L * L
        per_T = {}
        for c in range(n):
            T = float(Tc[c])
            Z, nbt, step, nbf = (int(v) for v in out[c, :4])
            if self._verbose:
                print("Running L = {}, T = {}, num_steps: {}".format(L, T, self._num_steps))
            if self._tree is not None:
                k = min(int(n_ev[c]), max_events)
                self._tree.write_chain(T, Z, nbt, step, nbf, events[c, :k],
                                       dumps[c, :k] if dumps is not None else None,
                                       final[c] if final is not None else None)
            line = compat_io.output_line(L, T, Z, nbt, step).split()
            # what np.loadtxt(output.txt) hands to worm_simulation.py:119-125
            vals = [float(x) for x in line]
            per_T.setdefault(temperature_key(vals[1]), []).append(vals)
        for key, rows in per_T.items():
            a = np.mean(np.array(rows), axis=0) if len(rows) > 1 else np.array(rows[0])
            self._beta[key], self._Z[key], self._E[key], self._Nb[key], self._sim_steps[key] = a[2], a[3], a[4], a[5], a[6]
            self._T.append(rows[0][1])

    # -- production mode ------------------------------------------------------------------------
    def _run_with_dump_cap(self, st, therm, max_dumps):
        """The executable dumps at every closed multiple of write_steps (worm_ising_2d.cpp:113-125); here the
        first `max_dumps` per chain are kept.  Once every chain of the shard has them, the dump rule is switched
        off for the rest of the run: the trajectory does not depend on it (chunked runs are bit-identical), and
        a segment boundary every write_steps steps costs about a third of the step rate (tools/dump_time.py)."""
        dump_every = self._write_steps if max_dumps else 0
        left = int(self._num_steps)
        if not dump_every:
            st.run(left, therm_steps=therm, dump_every=0, lanes_per_chain=self._lanes)
            return
        chunk = max(therm + 8 * dump_every * max_dumps, 50000)     # warm-up + ~8x the expected time to fill
        while left > 0:
            this = min(chunk, left)
            st.run(this, therm_steps=therm, dump_every=dump_every, lanes_per_chain=self._lanes)
            left -= this
            if left > 0 and int(st.n_dumps.min()) >= max_dumps:
                st.run(left, therm_steps=therm, dump_every=0, lanes_per_chain=self._lanes)
                left = 0
            chunk *= 2

    def _run_philox(self):
        import torch
        from . import dist
        L = self._L
        N = L * L
        Tc, r_idx, t_idx = self._chain_table()
        n_all = Tc.size
        nT = len(self._T_range)
        if self._seeds is None:
            seeds = r_idx.astype(np.uint64)
        else:
            seeds = np.asarray(self._seeds, dtype=np.uint64).reshape(-1)
            if seeds.size == self._n_replicas and self._n_replicas != n_all:
                seeds = seeds[r_idx]
            elif seeds.size == nT and nT != n_all:
                seeds = seeds[t_idx] + np.uint64(1000003) * r_idx.astype(np.uint64)
            if seeds.size != n_all:
                raise ValueError("seeds: one per replica, per temperature, or per chain")
        # Shard: contiguous block of the global chain range per rank.  Chain ids are global (they are
        # the Philox counter), so the reduced sums do not depend on the number of ranks.
        c0, c1 = dist.shard_range(n_all, self._rank, self._world)
        if self._world > n_all:
            raise ValueError("world_size {} exceeds the {} chains of the sweep".format(self._world, n_all))
        # the reference reads T back from "%.12f" text (worm_simulation.py:53,119-121): same keys, same _T
        T_round = np.array([float("%.12f" % t) for t in self._T_range])
        therm = int(self._therm_frac * self._num_steps)
        max_dumps = self._max_dumps
        if max_dumps is None:
            # dumps (2N bytes each) are kept for the first 64 events of a chain; without dumps the events alone
            # (observables rows, 24 bytes) are only wanted for the files
            max_dumps = 64 if self._bond_flag else (256 if self._tree is not None else 0)
        st = engine.PhiloxState(L, Tc[c0:c1], seeds[c0:c1], algo=self._algo, chain_id0=c0,
                                group_index=t_idx[c0:c1], n_groups=nT, hist=self._hist,
                                max_dumps=max_dumps, store_dumps=bool(self._bond_flag), device=self._device)
        self._run_with_dump_cap(st, therm, max_dumps)
        # per-temperature sums over the chains of this shard (device-side atomics), then over the ranks: a clone
        # is reduced, the state keeps its local sums and can be run on
        per_T = dist.allreduce_sum(st.group_acc.clone(), self._world)
        hist = None
        if self._hist:
            # G(r) histogram per temperature, [nT, N] with bin dy*L + dx; bin 0 == Z
            hist = dist.allreduce_sum(st.hist.clone(), self._world)
        acc_all = None
        if self._gather:
            # per-chain sums of every rank on every rank (the optional all-gather for error bars
            # over replicas): rows in global chain order
            acc_all = dist.allgather_rows(st.acc, n_all, self._rank, self._world)
        status = int((st.acc[:, ACC_STATUS] != 0).sum().item())
        if status:
            raise RuntimeError("{} chains report a bond occupation the packed kernel cannot hold (WORM_ACC_STATUS); "
                               "rerun with lanes_per_chain=1".format(status))
        self.results = dict(T=Tc[c0:c1], seeds=seeds[c0:c1], replica=r_idx[c0:c1], t_index=t_idx[c0:c1],
                            chain_range=(c0, c1), acc=st.acc, acc_all=acc_all, state=st,
                            dumps=st.dumps, n_dumps=st.n_dumps, events=st.events, bonds=st.bonds,
                            per_T=per_T, hist=hist, mode="philox", therm=therm)
        if self._tree is not None:
            self._write_tree_philox(st, Tc[c0:c1], therm, max_dumps)
        a = per_T.cpu().numpy().astype(np.float64)
        for ti_, T in enumerate(T_round):
            Z = a[ti_, ACC_Z]
            key = temperature_key(T)
            self._T.append(float(T))
            self._beta[key] = 1.0 / float(T)
            self._Z[key] = Z / max(a[ti_, ACC_ITERS], 1.0)
            self._E[key] = -float(T) * (a[ti_, ACC_NB] / max(Z, 1.0)) / N
            self._Nb[key] = (a[ti_, ACC_NB] / max(Z, 1.0)) / N
            self._sim_steps[key] = float(self._num_steps)
            if self._verbose:
                print("Ran L = {}, T = {}, num_steps: {} x {} replicas".format(L, T, self._num_steps, self._n_replicas))

    def _write_tree_philox(self, st, Tc, therm, max_dumps):
        """The reference's files from the device buffers of a production run: per chain the
        observables rows of its stored events (worm_ising_2d.cpp:113-119), with bond_flag its stored dumps and the
        final configuration (:120-124, :160-167), output.txt (:189-191) and the num_bonds line (:195-198)."""
        L = self._L
        acc = st.acc.cpu().numpy()
        nd = st.n_dumps.cpu().numpy()
        ev = st.events.cpu().numpy()
        want = bool(self._bond_flag)
        dumps = st.dumps.cpu().numpy() if (want and st.dumps is not None) else None
        final = st.bonds.cpu().numpy()
        nb_final = final.sum(axis=1, dtype=np.int64)
        for c in range(Tc.size):
            T = float("%.12f" % Tc[c])
            k = min(int(nd[c]), max_dumps)
            self._tree.write_chain_production(T, int(acc[c, ACC_Z]), int(acc[c, ACC_NB]), int(st.step), therm,
                                              int(nb_final[c]), ev[c, :k], dumps[c, :k] if dumps is not None else None,
                                              final[c] if want else None)
