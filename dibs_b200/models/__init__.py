from .graph import ErdosReniDAGDistribution, ScaleFreeDAGDistribution, UniformDAGDistributionRejection
from .linearGaussian import BGe, LinearGaussian
from .nonlinearGaussian import DenseNonlinearGaussian
