"""dibs_b200 -- B200-native (sm_100a) SVGD particle-update hot path of DiBS behind the reference's API.

    from dibs_b200.inference import JointDiBS, MarginalDiBS
    from dibs_b200.models import BGe, LinearGaussian, DenseNonlinearGaussian, ErdosReniDAGDistribution
    from dibs_b200.kernel import AdditiveFrobeniusSEKernel, JointAdditiveFrobeniusSEKernel
    from dibs_b200.target import make_linear_gaussian_model, make_nonlinear_gaussian_model      # data factory (host)
    from dibs_b200.metrics import expected_shd, threshold_metrics, neg_ave_log_likelihood       # evaluation (host)

The compute path is hand-written CUDA behind the C ABI of ``include/dibs_b200.h`` (bound with ctypes in
``dibs_b200/_native.py``); PyTorch only carries device memory, streams and ``torch.distributed``.
"""
__version__ = "0.1.0"
