"""gravity2_b200 -- B200-native (sm_100a) pairwise genome-similarity engine for GRAViTy-V2.

Importing the package touches neither CUDA nor the shared library; both are bound on first use.
"""
from .dropin import (  # noqa: F401
    SimilarityMat_Constructor, dist_mat_constructor, GOMDB_Constructor, GOMSignatureTable_Constructor,
    generate_gom_sigs, dcor, shared_pphmm_ratio, shared_norm_pphmm_ratio, shared_norm_pphmm_ratio_from_array,
    shared_pphmm_counts, binary_jaccard, bootstrap_distance_matrices, draw_resample_indices,
    weights_from_indices, gom_membership, DistMat2Tree, bootstrap_newick, sort_pphmm_distances, sort_pphmm_tree, braycurtis_matrix, pphmm_loc_diffs_pairwise,
    VirusGrouping_Estimator, theils_u, PairwiseSimilarityScore_Cutoff_Dict_Constructor, max_similarity, bootstrap_classification, load_annotator_tables,
)
from .engine import Engine, default_engine  # noqa: F401
from .multi import MultiEngine, MultiJob  # noqa: F401

__version__ = "0.1.0"
