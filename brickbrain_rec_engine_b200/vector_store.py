"""Vector-store seam: a LangChain-shaped flat store and the ``_query_vector_database``
replacement, both over ``ScoringEngine``.

Reference call sites this serves:

* ``PGVector.from_documents(documents=, embedding=, ...)`` / ``Chroma.from_documents`` and
  ``vectorstore.as_retriever(search_kwargs={"k": 20}).invoke(query)``
  (lego_nlp_recommeder.py:296-305, :315-320, :1394) -- third-party stores whose
  arithmetic is an exact flat cosine search over MiniLM (384-d, unit-norm) rows.
* the FAISS ``IndexFlatL2`` artefact (test_embeddings/index.faiss, ``save_local`` /
  ``load_local`` at lego_nlp_recommeder.py:1600-1654): ``read_faiss_flat`` parses it.
* ``HuggingFaceNLPRecommender._query_vector_database(query_embedding, top_k)``
  (hf_nlp_recommender.py:1310-1349), which in the reference is a stub returning random rows.
"""
from __future__ import annotations

import logging
import struct
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine import ScoringEngine

logger = logging.getLogger(__name__)

try:  # use LangChain's Document when it is installed, else a structural twin
    from langchain_core.documents import Document  # type: ignore
except Exception:  # pragma: no cover - langchain is not in this image
    class Document:
        def __init__(self, page_content: str, metadata: Optional[dict] = None):
            self.page_content = page_content
            self.metadata = metadata or {}

        def __repr__(self):
            return f"Document(page_content={self.page_content!r}, metadata={self.metadata!r})"


def read_faiss_flat(path: str) -> Tuple[np.ndarray, str]:
    """Parses a FAISS flat index file (fourcc ``IxF2`` = IndexFlatL2, ``IxFI`` = IndexFlatIP):
    header = fourcc, d:int32, ntotal:int64, 2 dummy int64, is_trained:u8, metric:int32, then a
    size-prefixed float32 vector of ntotal*d values.  Returns (rows [n, d] float32, metric)."""
    raw = open(path, "rb").read()
    fourcc = raw[:4]
    if fourcc not in (b"IxF2", b"IxFI"):
        raise ValueError(f"not a FAISS flat index: fourcc {fourcc!r}")
    d, = struct.unpack_from("<i", raw, 4)
    ntotal, = struct.unpack_from("<q", raw, 8)
    metric, = struct.unpack_from("<i", raw, 33)
    count, = struct.unpack_from("<q", raw, 37)
    if count != ntotal * d:
        raise ValueError("corrupt FAISS flat index: vector length mismatch")
    rows = np.frombuffer(raw, dtype="<f4", count=count, offset=45).reshape(ntotal, d).copy()
    return rows, ("l2" if fourcc == b"IxF2" else "ip")


class _Retriever:
    def __init__(self, store: "BrickVectorStore", k: int):
        self.store = store
        self.k = k

    def invoke(self, query: str) -> List[Document]:
        return self.store.similarity_search(query, k=self.k)

    get_relevant_documents = invoke


class BrickVectorStore:
    """Exact flat vector store with the subset of the LangChain ``VectorStore``
    contract the reference uses.  ``score_mode='cosine'`` ranks like pgvector's
    ``<=>``; ``'l2'`` reports FAISS IndexFlatL2 distances (2 - 2cos on unit rows)."""

    def __init__(self, embeddings: np.ndarray, documents: Sequence[Document],
                 embedding_function: Optional[Any] = None, score_mode: str = "cosine", device: int = 0,
                 precision: str = "fp32", k_max: int = 128):
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        if emb.shape[0] != len(documents):
            raise ValueError("one embedding per document required")
        self.documents = list(documents)
        self.embedding_function = embedding_function
        self.score_mode = score_mode
        self.precision = precision
        self.engine = ScoringEngine(n_items=emb.shape[0], dim=emb.shape[1], cf_dim=0, device=device, k_max=k_max)
        self.engine.load_catalog(emb=emb)

    # ---- constructors used by the reference ------------------------------------------------
    @classmethod
    def from_documents(cls, documents: Sequence[Document], embedding: Any, **kwargs) -> "BrickVectorStore":
        texts = [d.page_content for d in documents]
        vecs = np.asarray(embedding.embed_documents(texts), dtype=np.float32)
        keep = {k: v for k, v in kwargs.items() if k in ("score_mode", "device", "precision", "k_max")}
        return cls(vecs, documents, embedding_function=embedding, **keep)

    @classmethod
    def from_embeddings(cls, embeddings: np.ndarray, metadatas: Sequence[dict], texts: Optional[Sequence[str]] = None,
                        **kwargs) -> "BrickVectorStore":
        docs = [Document(page_content=(texts[i] if texts else ""), metadata=dict(m)) for i, m in enumerate(metadatas)]
        return cls(embeddings, docs, **kwargs)

    @classmethod
    def load_local(cls, faiss_path: str, documents: Sequence[Document], **kwargs) -> "BrickVectorStore":
        rows, metric = read_faiss_flat(faiss_path)
        kwargs.setdefault("score_mode", "l2" if metric == "l2" else "cosine")
        return cls(rows, documents, **kwargs)

    # ---- search -----------------------------------------------------------------------------
    def _search(self, vectors: np.ndarray, k: int):
        q = np.ascontiguousarray(np.atleast_2d(vectors), dtype=np.float32)
        k_eff = min(k, len(self.documents), self.engine.k_max)
        scores, ids = self.engine.search(k_eff, q_emb=q, w_content=1.0, w_cf=0.0, precision=self.precision,
                                         as_l2=self.score_mode == "l2")
        return scores, ids

    def similarity_search_with_score_by_vector(self, embedding: Sequence[float], k: int = 4
                                               ) -> List[Tuple[Document, float]]:
        scores, ids = self._search(np.asarray(embedding, dtype=np.float32), k)
        return [(self.documents[int(i)], float(s)) for i, s in zip(ids[0], scores[0]) if i >= 0]

    def similarity_search_by_vector(self, embedding: Sequence[float], k: int = 4) -> List[Document]:
        return [d for d, _ in self.similarity_search_with_score_by_vector(embedding, k)]

    def _embed_query(self, query: str) -> np.ndarray:
        if self.embedding_function is None:
            raise RuntimeError("this store was built from raw vectors; pass embedding_function to search by text")
        return np.asarray(self.embedding_function.embed_query(query), dtype=np.float32)

    def similarity_search(self, query: str, k: int = 4) -> List[Document]:
        return self.similarity_search_by_vector(self._embed_query(query), k)

    def similarity_search_with_score(self, query: str, k: int = 4) -> List[Tuple[Document, float]]:
        return self.similarity_search_with_score_by_vector(self._embed_query(query), k)

    def similarity_search_batch(self, embeddings: np.ndarray, k: int = 4):
        """Batched form (no LangChain counterpart): (scores [B,k], ids [B,k])."""
        return self._search(embeddings, k)

    def as_retriever(self, search_kwargs: Optional[dict] = None) -> _Retriever:
        return _Retriever(self, int((search_kwargs or {}).get("k", 4)))


def make_query_vector_database(store: BrickVectorStore) -> Callable[[np.ndarray, int], List[Dict]]:
    """Returns a function with the signature and result schema of
    ``HuggingFaceNLPRecommender._query_vector_database`` (hf_nlp_recommender.py:1310-1349):
    ``(query_embedding: np.ndarray[384], top_k: int) -> List[Dict]`` with keys
    set_num, name, year, num_parts, theme, theme_id, img_url, relevance_score.  Errors are logged
    and produce ``[]`` like the reference (:1347-1349).  Bind it with
    ``recommender._query_vector_database = make_query_vector_database(store)``."""

    def _query_vector_database(query_embedding: np.ndarray, top_k: int) -> List[Dict]:
        try:
            out = []
            for doc, score in store.similarity_search_with_score_by_vector(query_embedding, top_k):
                m = doc.metadata
                out.append({"set_num": m.get("set_num"), "name": m.get("name"), "year": m.get("year"),
                            "num_parts": m.get("num_parts"), "theme": m.get("theme") or "Generic",
                            "theme_id": m.get("theme_id"), "img_url": m.get("img_url"),
                            "relevance_score": score})
            return out[:top_k]
        except Exception as e:
            logger.error("Database query failed: %s", e)
            return []

    return _query_vector_database


# ---------------------------------------------------------------------------------------------
# the callers' post-filters (what happens to the store's hits before they are returned)
# ---------------------------------------------------------------------------------------------
def _document_rules(filters: Dict[str, Any]):
    """The legacy recommender's post-filter (lego_nlp_recommeder.py:1514-1549) as a list of predicates over a document's
    metadata; a falsy filter value means "not set", as in the reference."""
    f = filters or {}
    rules = []
    if f.get("min_pieces"):
        rules.append(lambda m, v=f["min_pieces"]: m["num_parts"] >= v)
    if f.get("max_pieces"):
        rules.append(lambda m, v=f["max_pieces"]: m["num_parts"] <= v)
    if f.get("themes"):
        wanted = [t.lower() for t in f["themes"]]
        rules.append(lambda m: any(t in m["theme"].lower() for t in wanted))          # substring, case-insensitive
    if f.get("complexity"):
        rules.append(lambda m, v=f["complexity"]: m.get("complexity") == v)
    if f.get("min_age"):
        rules.append(lambda m: m["year"] >= 2010)                                     # "newer sets for younger builders"
    return rules


def apply_document_filters(docs: Sequence[Document], filters: Optional[Dict[str, Any]]) -> List[Document]:
    """``NLPRecommender._apply_filters``: keeps the documents whose metadata passes every active rule, in order."""
    rules = _document_rules(filters)
    return [d for d in docs if all(rule(d.metadata) for rule in rules)]


_RESULT_RULES = (          # (filter key, result key, keep-if) of hf_nlp_recommender.py:1351-1386
    ("themes", "theme", lambda have, want: have in want),
    ("min_pieces", "num_parts", lambda have, want: have >= want),
    ("max_pieces", "num_parts", lambda have, want: have <= want),
    ("min_age", "min_age", lambda have, want: have <= want),
    ("max_age", "max_age", lambda have, want: have >= want),
)


def apply_result_filters(results: Sequence[Dict], filters: Optional[Dict[str, Any]]) -> List[Dict]:
    """``HuggingFaceNLPRecommender._apply_filters``: a rule only applies when BOTH the filter value and the result's
    field are truthy (a hit without ``min_age`` passes the age filter); themes match by exact membership."""
    if not filters:
        return list(results)
    active = [(rk, keep, filters[fk]) for fk, rk, keep in _RESULT_RULES if filters.get(fk)]
    return [r for r in results if all(not r.get(rk) or keep(r[rk], want) for rk, keep, want in active)]


def semantic_search(store: BrickVectorStore, query, top_k: int = 10, filters: Optional[Dict[str, Any]] = None,
                    retrieve_k: int = 20) -> List[Dict]:
    """The search half of ``NLPRecommender.semantic_search`` (lego_nlp_recommeder.py:1394-1411): the retriever's best
    ``retrieve_k`` documents (k=20, :305) -> post-filter -> ``[:top_k]``.  Like the reference, nothing is re-fetched
    when the filter removes hits, so fewer than ``top_k`` results may come back.  ``query``: text (needs the store's
    embedding function) or an embedding vector."""
    if isinstance(query, str):
        hits = store.similarity_search_with_score(query, k=retrieve_k)
    else:
        hits = store.similarity_search_with_score_by_vector(query, k=retrieve_k)
    kept = apply_document_filters([d for d, _ in hits], filters)
    score_of = {id(d): s for d, s in hits}
    return [{"set_num": d.metadata["set_num"], "name": d.metadata["name"], "year": d.metadata["year"],
             "num_parts": d.metadata["num_parts"], "theme": d.metadata["theme"], "description": d.page_content,
             "score": score_of[id(d)]} for d in kept[:top_k]]
