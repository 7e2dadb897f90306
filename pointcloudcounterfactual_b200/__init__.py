"""pointcloudcounterfactual_b200 -- B200 (sm_100a) geometry hot path for nverchev/PointCloudCounterfactual.

Operators (all CUDA-only, hand-written kernels behind the C ABI of ``include/pcc_b200.h``):
    structural_losses.nn_distance / match_cost      Chamfer NN distance, approximate-matching EMD
    emd.emdModule                                   auction EMD
    neighbour_ops.knn, get_graph_features, ...      DGCNN kNN graph construction
    keops.LazyTensor                                the four KeOps patterns the reference uses
    losses.pykeops_chamfer / torch_chamfer          the reference's Chamfer reductions
    sharding                                        batch sharding over ranks + NCCL loss reduction

Importing the package does not load the CUDA library (so host-side logic is testable without a GPU); the first
operator call does, and raises if it is missing.  There is no CPU fallback.
"""
__version__ = "0.1.0"
