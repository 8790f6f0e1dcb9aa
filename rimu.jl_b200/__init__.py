"""rimu.jl_b200 -- B200-native FCIQMC step behind the Rimu.jl API.

Python host mirror of the reference's operator/vector interface for the hot path
(ProjectorMonteCarloProblem/solve, the Hamiltonian interface, Fock addresses, DVec/PDVec and the
StochasticStyles) over a C-ABI CUDA library (include/rimu_b200.h).  Import as `rimu_b200`
(root shim) -- the directory name contains a dot and cannot be imported directly.
"""
from . import _lib
from ._lib import RimuB200Error, build
from .addresses import (AddressType, BoseFS, CompositeFS, FermiFS, FermiFS2C, near_uniform, near_uniform_onr,
                        num_modes, num_particles, onr)
from .hamiltonians import (AbstractHamiltonian, Context, CubicGrid, ExtendedHubbardMom1D, ExtendedHubbardReal1D, HubbardMom1DEP, HardwallBoundaries, HubbardMom1D,
                           HubbardReal1D, HubbardReal1DEP, shift_lattice, shift_lattice_inv,
                           HubbardRealSpace, LadderBoundaries, PeriodicBoundaries, Transcorrelated1D,
                           continuum_dispersion, diagonal_element, dimension, get_context, get_offdiagonal,
                           hubbard_dispersion, num_offdiagonals, offdiagonals, random_offdiagonal, reset_contexts,
                           starting_address)
from .stochasticstyles import (IsDeterministic, IsDynamicSemistochastic, IsStochasticInteger,
                               IsStochasticWithThreshold, NoCompression, StochasticStyle, ThresholdCompression,
                               default_style, step_stats,
                               CoherentInitiator, Initiator, InitiatorRule, NonInitiator, SimpleInitiator)
from .dictvectors import (DVec, FirstOrderTransitionOperator, FrozenDVec, GPUDVec, InitiatorDVec, PDVec, WorkingMemory, advance, apply_operator, dot, mul,
                          walkernumber_and_length, working_memory)
from .fciqmc import (AllOverlaps, DataFrame, DontUpdate, DoubleLogUpdate, DoubleLogUpdateAfterTargetWalkers, GramSchmidt, LogUpdate, ReportDFAndInfo, ReportToFile, load_df,
                     LogUpdateAfterTargetWalkers,
                     PMCSimulation, ProjectedEnergy, Projector, ProjectorMonteCarloProblem, ShiftParameters,
                     SingleState, Timer, default_starting_vector, init, solve, solve_, step_)
from .statstools import blocking_analysis, projected_energy, ratio_of_means, shift_estimator
from .lanczos import eigsolve_lanczos
from .sectors import DenseSectorVec, SectorBasis
from .rimuio import load_state, save_state
from .distributed import init_distributed
