"""Pair interactions: the reference's ``rock_paper_scissors`` factory, dispatched to CUDA.

Mirrors /root/reference/interactions.py.  ``rock_paper_scissors(N, pRS, pPR, pSP)`` returns the
same triple ``(interaction, parameters, {"species": int8[N]})`` (interactions.py:43-54); the first
element is a marker that ``InteractionSimulator`` recognises and runs on the device
(csrc/pairs.cu, csrc/resolve.cu).  The rule itself (interactions.py:13-40) is not re-implemented
on the host: calling the marker raises, because this package has no CPU path.
"""
import numpy as np
from numpy import int8

# interactions.py:5 -- 8-bit species codes
ROCK, PAPER, SCISSORS = int8(1), int8(2), int8(3)

# interactions.py:8-10 -- plotting colours
ROCK_COLOR = "red"
PAPER_COLOR = "limegreen"
SCISSORS_COLOR = "blue"


class DeviceInteraction:
    """Tag for an interaction rule that exists as a CUDA kernel."""

    def __init__(self, name):
        self.__name__ = name

    def __call__(self, parameters, microbe_properties, p1, p2):
        raise NotImplementedError(
            "%s runs on the GPU inside InteractionSimulator.time_step; it cannot be called per pair on the host"
            % self.__name__)

    def __repr__(self):
        return "<device interaction %s>" % self.__name__


rock_paper_scissors_interaction = DeviceInteraction("rock_paper_scissors_interaction")


def is_rock_paper_scissors(fn):
    """True for our marker and for the reference's own function object (recognised by name)."""
    return getattr(fn, "__name__", None) == "rock_paper_scissors_interaction"


def rock_paper_scissors(N_microbes, pRS, pPR, pSP):
    microbe_properties = {
        "species": np.random.choice([ROCK, PAPER, SCISSORS], N_microbes)      # int8[N], interactions.py:45
    }
    interaction_parameters = {
        "pRS": pRS,  # forward probability that rock beats scissors
        "pPR": pPR,  # forward probability that paper beats rock
        "pSP": pSP,  # forward probability that scissors beats paper
    }
    return rock_paper_scissors_interaction, interaction_parameters, microbe_properties
