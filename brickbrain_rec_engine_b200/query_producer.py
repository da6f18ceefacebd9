"""The step immediately BEFORE the scoring path: the query-embedding producer.

Reference: ``HuggingFaceNLPRecommender.search_recommendations`` encodes the semantic query with MiniLM
(``self.embedding_model.encode([semantic_query])``, hf_nlp_recommender.py:1242; the legacy path builds the same
384-d, ``normalize_embeddings=True`` vectors, lego_nlp_recommeder.py:143-149), takes the NumPy row, and calls
``_query_vector_database`` (:1245) -- one host round trip per query, one query per call.

Here the encoder's output never leaves the GPU: ``DeviceQueryProducer`` accepts any ``encode`` callable that returns a
CUDA tensor [n, dim] (``SentenceTransformer.encode(..., convert_to_tensor=True)``, a torch module, ...) and hands the
tensor's device pointer to ``bb_search`` on the SAME stream, so encoder and scan are ordered by the stream and the host
only enqueues.  Queries can be accumulated (``submit`` / ``flush``) so that many requests share one scan of the catalog:
below ~255 queries the scan is HBM-bound (SURVEY 8d ridge), i.e. the 2nd..255th query of a batch is almost free.
"""
from __future__ import annotations

import logging
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from .vector_store import BrickVectorStore, apply_result_filters

logger = logging.getLogger(__name__)

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class DeviceQueryProducer:
    """``encode(texts) -> CUDA tensor [n, dim]`` feeding ``store``'s engine without a host hop."""

    def __init__(self, encode: Callable[[Sequence[str]], Any], store: BrickVectorStore, max_pending: int = 256):
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("DeviceQueryProducer needs a CUDA device (there is no CPU fallback)")
        self.encode = encode
        self.store = store
        self.max_pending = max_pending
        self._pending: List[str] = []
        self._tickets: List[int] = []
        self._results: Dict[int, Any] = {}
        self._next_ticket = 0

    # ---- one call, any number of queries ---------------------------------------------------------------
    def search_tensor(self, texts_or_embeddings, k: int, precision: Optional[str] = None):
        """(scores [n, k], ids [n, k]) as CUDA tensors, enqueued on the current stream (no synchronization)."""
        q = texts_or_embeddings
        if not torch.is_tensor(q):
            q = self.encode(list(q))
        if not (torch.is_tensor(q) and q.is_cuda):
            raise TypeError("encode() must return a CUDA tensor: the producer exists to keep embeddings on the device")
        if q.dim() == 1:
            q = q.unsqueeze(0)
        eng = self.store.engine
        k_eff = min(k, len(self.store.documents), eng.k_max)
        return eng.search(k_eff, q_emb=q.to(torch.float32), w_content=1.0, w_cf=0.0,
                          precision=precision or self.store.precision, as_l2=self.store.score_mode == "l2")

    # ---- micro-batching ----------------------------------------------------------------------------------
    def submit(self, text: str) -> int:
        """Queues one query; returns a ticket for ``result``.  The batch runs at ``flush`` or when it is full."""
        t = self._next_ticket
        self._next_ticket += 1
        self._pending.append(text)
        self._tickets.append(t)
        if len(self._pending) >= self.max_pending:
            self.flush()
        return t

    def flush(self, k: int = 20):
        if not self._pending:
            return
        scores, ids = self.search_tensor(self._pending, k)
        for row, t in enumerate(self._tickets):
            self._results[t] = (scores[row], ids[row])
        self._pending, self._tickets = [], []

    def result(self, ticket: int, k: int = 20):
        if ticket not in self._results:
            self.flush(k)
        return self._results.pop(ticket)

    # ---- the reference's seam ------------------------------------------------------------------------------
    def search_recommendations(self, processed_query: Dict[str, Any], top_k: int = 10) -> List[Dict[str, Any]]:
        """The general-search branch of ``HuggingFaceNLPRecommender.search_recommendations`` (hf_nlp_recommender.py:1240-1259;
        the theme shortcut before it is a SQL query and stays with the caller): encode the semantic query, search, apply
        the result filters (:1351-1386), attach confidence / intent, ``[:top_k]``.  Errors are logged and give ``[]``."""
        try:
            scores, ids = self.search_tensor([processed_query["semantic_query"]], top_k)
            scores, ids = scores[0].cpu().numpy(), ids[0].cpu().numpy()       # the ONLY device->host copy: k (score, id) pairs
            results = []
            for s, i in zip(scores, ids):
                if i < 0:
                    break
                m = self.store.documents[int(i)].metadata
                results.append({"set_num": m.get("set_num"), "name": m.get("name"), "year": m.get("year"),
                                "num_parts": m.get("num_parts"), "theme": m.get("theme") or "Generic",
                                "theme_id": m.get("theme_id"), "img_url": m.get("img_url"), "relevance_score": float(s)})
            filtered = apply_result_filters(results, processed_query.get("filters", {}))
            for r in filtered:
                r["confidence"] = processed_query.get("confidence")
                r["intent"] = processed_query.get("intent")
            return filtered[:top_k]
        except Exception as e:
            logger.error("Search failed: %s", e)
            return []


def minilm_shaped_encoder(device, vocab_size: int = 30522, seed: int = 0):
    """A random-init encoder with the architecture of all-MiniLM-L6-v2 (6 layers, hidden 384, 12 heads, mean pooling,
    L2-normalized output -- lego_nlp_recommeder.py:143-149) for tests and benchmarks: there are no checkpoints in this
    image.  Returns ``encode(list of token-id lists | list of str) -> CUDA tensor [n, 384]``.  Strings are hashed into
    token ids (no tokenizer vocabulary is available offline)."""
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=vocab_size, hidden_size=384, num_hidden_layers=6, num_attention_heads=12,
                     intermediate_size=1536, max_position_embeddings=512)
    model = BertModel(cfg, add_pooling_layer=False).to(device).eval().half()

    def encode(texts: Sequence) -> "torch.Tensor":
        rows = []
        for t in texts:
            if isinstance(t, str):
                t = [101] + [1000 + (hash(w) % (vocab_size - 1000)) for w in t.lower().split()][:126] + [102]
            rows.append(list(t))
        n = max(len(r) for r in rows)
        ids = torch.zeros((len(rows), n), dtype=torch.long)
        mask = torch.zeros((len(rows), n), dtype=torch.long)
        for j, r in enumerate(rows):
            ids[j, :len(r)] = torch.tensor(r)
            mask[j, :len(r)] = 1
        ids, mask = ids.to(device, non_blocking=True), mask.to(device, non_blocking=True)
        with torch.no_grad():
            h = model(input_ids=ids, attention_mask=mask).last_hidden_state.float()
            m = mask.unsqueeze(-1).float()
            pooled = (h * m).sum(1) / m.sum(1).clamp(min=1.0)
            return torch.nn.functional.normalize(pooled, dim=1)

    return encode
