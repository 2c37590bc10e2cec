"""waves.jl_b200 -- B200-native drop-in for the Waves.jl acoustic RK4 hot path.

The computation lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of include/waves_b200.h);
this package is the host-side mirror of the reference's call surface for that path.
"""
from . import _lib, bson
from ._lib import WavesError, build
from .engine import ADJ_COMPAT, ADJ_EXACT, MODE_EXACT, MODE_FUSED, STEP_ASYNC, Engine
from .env import AcousticDynamics, Integrator, WaveEnv
from .data import (BatchWaveEnv, Episode, RandomDesignPolicy, flatten_repeated_last_dim, generate_episode, generate_episodes,
                   prepare_data)
from .latent import LATENT_ADJ_R1, LATENT_AUTO, LATENT_GENERIC, LATENT_PAIR, LATENT_SINGLE, LatentDynamics, LatentIntegrator, LatentSource, LinearInterpolation, OneDim, build_pml_1d
from .parallel import HaloExchanger, LocalSlabGroup, SlabEngine, shard_envs, slab_rows
from .host import (AIR, WATER, Cloak, Cylinders, DesignInterpolator, DesignSpace, NoSource, RandomPosGaussianSource,
                   Source, TwoDim, build_action_space, build_gradient, build_grid, build_normal, build_pml,
                   build_radii_design_space, build_triple_ring_design_space, build_tspan, build_wave, get_dx, get_dy,
                   hexagon_ring, julia_range)

__all__ = [n for n in dir() if not n.startswith("_")]
