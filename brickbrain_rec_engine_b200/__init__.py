"""brickbrain_rec_engine_b200 -- the B200-native candidate-scoring path of
Brickbrain-Rec-Engine (content cosine, CF dot, hybrid blend, hard-constraint mask,
top-k) behind the reference's own Python entry points.

The arithmetic lives in ``libbrickscore.so`` (hand-written sm_100a CUDA, C ABI in
``include/brickscore.h``); this package is the host-side mirror of the reference
interface.  There is no CPU fallback: importing works without a GPU, computing does not.
"""
from ._lib import BrickscoreError, CONSTRAINT_DTYPE, no_constraint  # noqa: F401
from .engine import ScoringEngine  # noqa: F401
from .constraints import (ConstraintCompiler, ConstraintResult, ConstraintSeverity, ConstraintType,  # noqa: F401
                          ConstraintViolation, HardConstraint, HardConstraintFilter, theme_bits_for,
                          create_age_appropriate_constraints, create_budget_constraints, create_size_constraints)
from .recommenders import (CatalogTable, CollaborativeFilteringRecommender, ContentBasedRecommender,  # noqa: F401
                           HybridRecommender, RecommendationRequest, RecommendationResult, accelerate)
from .vector_store import (BrickVectorStore, Document, apply_document_filters, apply_result_filters,  # noqa: F401
                           make_query_vector_database, read_faiss_flat, semantic_search)
from .sharded import ShardedSearch, shard_bounds, slice_bounds  # noqa: F401
from .bulk import BulkSink, bulk_recommend, user_shard  # noqa: F401
from .cache import RecommendationCache  # noqa: F401
from .query_producer import DeviceQueryProducer, minilm_shaped_encoder  # noqa: F401

__version__ = "0.1.0"
