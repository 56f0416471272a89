"""lagrangian_microbes_b200 -- the B200-native per-timestep hot path of lagrangian-microbes.

RK4 advection, radius pair search and rock-paper-scissors resolution as hand-written sm_100a CUDA
(csrc/, C ABI in include/lm_b200.h) behind the reference's own Python driver API:

    uniform_particle_locations, ParticleAdvecter      (reference: particle_advecter.py)
    rock_paper_scissors                                (reference: interactions.py)
    InteractionSimulator                               (reference: interaction_simulator.py)
    velocity_fields.oscar_dataset                      (reference: velocity_fields.py)
    microbe_plotter.MicrobePlotter                     (reference: microbe_plotter.py; frames rasterised on the GPU)
    analysis.species_count_figure, analysis.pairwise_distance_histogram
                                                       (reference: analysis.py, sandbox/pairwise_distance_histogram_*.jl)

plus ``FusedSimulation``, the device-resident fused loop.  Importing this package does not touch
the GPU; constructing any compute object without the CUDA library / a CUDA device raises.
"""
from .interactions import (PAPER, ROCK, SCISSORS, rock_paper_scissors,  # noqa: F401
                           rock_paper_scissors_interaction)
from .particle_advecter import (OutOfBoundsError, ParticleAdvecter, TimeExtrapolationError,  # noqa: F401
                                distribute_particles_across_tiles, uniform_particle_locations)
from .interaction_simulator import InteractionSimulator  # noqa: F401
from . import velocity_fields  # noqa: F401
from . import analysis  # noqa: F401
from .microbe_plotter import MicrobePlotter  # noqa: F401

__all__ = ["ROCK", "PAPER", "SCISSORS", "rock_paper_scissors", "rock_paper_scissors_interaction",
           "ParticleAdvecter", "InteractionSimulator", "uniform_particle_locations",
           "distribute_particles_across_tiles", "OutOfBoundsError", "TimeExtrapolationError", "velocity_fields",
           "FusedSimulation", "MicrobePlotter", "analysis"]


def __getattr__(name):
    if name == "FusedSimulation":          # imports torch; keep the package import light
        from .simulation import FusedSimulation
        return FusedSimulation
    raise AttributeError(name)
