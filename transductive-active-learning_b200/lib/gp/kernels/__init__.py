from lib.gp.kernels.base import Kernel, Parameterized
import lib.gp.kernels.stationary as stationary
import lib.gp.kernels.non_stationary as non_stationary
