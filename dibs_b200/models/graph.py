"""Graph-prior plugins: plain descriptors the native plan consumes.

Mirror the constructors of the reference's graph models (dibs/models/graph.py:10-30, 111-130, 199-215).
Only what the SVGD step uses -- ``unnormalized_log_prob_soft`` and its gradient through the edge
probabilities (graph.py:93-108, 182-196, 263-276) -- is implemented, inside the CUDA assemble kernel
(dibs_b200/csrc/kernels_prior.cuh).  ``sample_G`` (ground-truth DAG sampling) is data generation and lives in
``dibs_b200.synthetic``.
"""


class ErdosReniDAGDistribution:
    """p(G) ~ p^e (1-p)^(C(d,2)-e); ``p`` set so that a node has ``n_edges_per_node`` edges in expectation."""
    native_kind = "er"

    def __init__(self, n_vars, n_edges_per_node=2):
        self.n_vars = n_vars
        self.n_edges = n_edges_per_node * n_vars
        self.p = self.n_edges / ((self.n_vars * (self.n_vars - 1)) / 2)


class ScaleFreeDAGDistribution:
    """p(G) ~ prod_j (1 + indegree_j)^-3."""
    native_kind = "sf"

    def __init__(self, n_vars, verbose=False, n_edges_per_node=2):
        self.n_vars = n_vars
        self.n_edges_per_node = n_edges_per_node
        self.verbose = verbose
        self.p = 0.0


class UniformDAGDistributionRejection:
    """Uniform prior over DAGs: contributes nothing to the latent prior score."""
    native_kind = "uniform"

    def __init__(self, n_vars):
        self.n_vars = n_vars
        self.p = 0.0
