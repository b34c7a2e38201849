"""rbm_b200 -- Python host mirror of the ReducedBasisMethods.jl API surface for the offline hot path,
calling librbm_b200.so (hand-written FP64 sm_100a CUDA) through its C ABI.  See DESIGN.md."""
from . import bump_on_tail as BumpOnTail
from ._lib import Context, RBMError, EXPORTS, LIB_PATH
from .engine import default_context, set_default_context
from .integrator import (VPIntegratorCache, VPIntegratorParameters, integrate_vp, integrate_vp_all)
from .parameter import CartesianParameterSampler, Parameter, ParameterSpace, named_parameters, sample
from .particles import ParticleList
from .poisson import B200PoissonField, PICHandle, PoissonSolverPBSplines, ScaledPoissonField
from .reducedbasis import (CotangentLiftEVD, CotangentLiftSVD, ReducedBasis, ReductionAlgorithm, UnspecifiedAlgorithm,
                           deim_indices, get_DEIM_interpolation_matrix, get_PODBasis_EVD,
                           get_PODBasis_cotangentLiftEVD, gram, project, sorteigen)
from .reduced import DEIMElectricField, ReducedModel, reduced_integrate_vp, reduced_integrate_vp_all
from .regression import find_maxima, get_regression_alpha_beta, get_regression_alpha_beta_device
from .snapshots import IntegratorParameters, Snapshots
from .trainingset import TrainingSet
