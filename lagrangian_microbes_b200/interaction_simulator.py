"""InteractionSimulator: the reference's interaction driver on the B200 pair-search + RPS kernels.

Mirrors /root/reference/interaction_simulator.py: same constructor (:23-52), same
``time_step(start_time, end_time, dt)`` (:54-122) reading ``particle_data.nc`` from
``advection_dir`` and writing ``microbe_data.nc`` to ``output_dir``.  The hot loop (:82-117)

    kdt = cKDTree(locations); pairs = kdt.query_pairs(r, p)        # :93, :98
    for pair in pairs: pair_interaction(params, props, *pair)       # :104-105

becomes one ``lm_interact_rps`` call per step (binning + fused pair search / RPS resolution in
the canonical cell-phase order).  Per-pair random numbers come from
Philox(seed, step=iteration, i, j) instead of NumPy's global stream (interactions.py:20).

Only the rock-paper-scissors triple can run on the device; any other ``pair_interaction`` callable
raises NotImplementedError (arbitrary Python cannot execute in a kernel, and there is no CPU path).
``interaction_norm`` may be 1, 2 (the only value the reference's scripts use) or ``math.inf`` -- the norms
SciPy evaluates without pow(), reproduced bit for bit (LM_OPT_NORM); ``self_interaction`` is accepted and
ignored exactly as in the reference (:28, :49).
"""
import logging
import os

import numpy as np
from numpy import float32, int8

from . import io as lmio
from .interactions import is_rock_paper_scissors
from .utils import pretty_time

logger = logging.getLogger(__name__)


class InteractionSimulator:
    def __init__(
            self,
            pair_interaction,
            interaction_radius,
            interaction_norm=2,
            self_interaction=None,
            advection_dir=".",
            output_dir=".",
            seed=0,
            pair_capacity=None,
    ):
        output_dir = os.path.abspath(output_dir)
        if not os.path.exists(output_dir):
            logger.info("Creating directory: {:s}".format(output_dir))
            os.makedirs(output_dir)

        pair_interaction_function, pair_interaction_parameters, microbe_properties = pair_interaction
        if not is_rock_paper_scissors(pair_interaction_function):
            raise NotImplementedError("only the rock_paper_scissors interaction exists as a CUDA kernel; "
                                      "arbitrary Python pair interactions cannot run on the device")
        from .engine import norm_code
        norm_code(interaction_norm)          # raises NotImplementedError for any p other than 1, 2, inf

        self.microbe_properties = microbe_properties
        self.pair_interaction = pair_interaction_function
        self.pair_interaction_parameters = pair_interaction_parameters
        self.interaction_radius = interaction_radius
        self.interaction_norm = interaction_norm
        self.self_interaction = self_interaction
        self.advection_dir = advection_dir
        self.output_dir = output_dir
        self.iteration = 0
        self.seed = seed
        self.pair_capacity = pair_capacity   # pairs per step the device buffers hold (default 32 per microbe)
        self.pairs_found = []          # per-step pair counts of the last time_step call
        self._engine = None

    def time_step(self, start_time, end_time, dt):
        import torch
        from .engine import Engine, make_grid

        particle_data = lmio.read_particle_file(os.path.join(self.advection_dir, "particle_data.nc"))
        lon_all, lat_all = particle_data["longitude"], particle_data["latitude"]
        N_particles, Nt = lon_all.shape
        times = [start_time + n * dt for n in range(Nt)]

        # The reference fills three dense (N, Nt) host arrays (:80-82) -- 33.8 GB for its own 490,000 x 7,670 run.  Same
        # file here, written column by column: in memory while it fits NetCDF-3, in blocks to memory-mapped files beyond.
        nc_output_filepath = os.path.join(self.output_dir, "microbe_data.nc")
        out = lmio.ParticleFileWriter(nc_output_filepath, {"longitude": float32, "latitude": float32, "species": int8},
                                      N_particles, times)

        if self._engine is None or self._engine.max_particles < N_particles:
            cap = self.pair_capacity if self.pair_capacity is not None else max(32 * N_particles, 1 << 20)
            self._engine = Engine(max_particles=N_particles, max_cells=max(4 * N_particles, 1 << 18), max_pairs=int(cap))
        eng = self._engine
        eng.set_norm(self.interaction_norm)
        dev = eng.device
        prm = self.pair_interaction_parameters
        r = float(self.interaction_radius)

        species_host = np.ascontiguousarray(self.microbe_properties["species"], dtype=np.int8)
        assert species_host.size == N_particles
        species = torch.from_numpy(species_host.copy()).to(dev)
        lon_pin = torch.empty(N_particles, dtype=torch.float32).pin_memory()
        lat_pin = torch.empty(N_particles, dtype=torch.float32).pin_memory()
        sp_pin = torch.empty(N_particles, dtype=torch.int8).pin_memory()
        lon_d = torch.empty(N_particles, dtype=torch.float32, device=dev)
        lat_d = torch.empty(N_particles, dtype=torch.float32, device=dev)

        logger.info("Simulating interactions for {:d} microbes over {:d} time steps ({:} -> {:})."
                    .format(N_particles, Nt, start_time, end_time))
        self.pairs_found = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = start_time
        while t < end_time:
            i = self.iteration
            lon_pin.numpy()[:] = lon_all[:, i]                                   # :88-89
            lat_pin.numpy()[:] = lat_all[:, i]
            lo, la = lon_pin.numpy(), lat_pin.numpy()
            eng.set_grid(make_grid(float(lo.min()), float(lo.max()), float(la.min()), float(la.max()), r,
                                   N_particles, eng.max_cells, margin=0.0))
            ev0.record()
            lon_d.copy_(lon_pin, non_blocking=True)
            lat_d.copy_(lat_pin, non_blocking=True)
            eng.interact_rps(lon_d, lat_d, species, r, prm["pRS"], prm["pPR"], prm["pSP"], self.seed, i)  # :93-105
            sp_pin.copy_(species, non_blocking=True)
            ev1.record()
            st = eng.sync_stats()
            self.pairs_found.append(int(st.n_pairs))
            logger.info("Step {:d}: {:d} pairs; bin + pair search + interactions: {:s}."
                        .format(i, st.n_pairs, pretty_time(ev0.elapsed_time(ev1) * 1e-3)))

            out.put(i, longitude=lo, latitude=la, species=sp_pin.numpy())       # :108-110
            t = t + dt
            self.iteration += 1

        # the reference mutates microbe_properties["species"] in place (interactions.py:37-40)
        if self.pairs_found:
            self.microbe_properties["species"][:] = sp_pin.numpy()
        logger.info("Writing microbe data to {:s}...".format(nc_output_filepath))
        out.close()
