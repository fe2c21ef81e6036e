"""B200-native query-time search scoring for anime-illust-image-searcher (the webui.py hot path).

Import as ``ais_b200`` (see ../ais_b200.py).  Sub-modules:
  synth      synthetic index / query generator (numpy; no GPU needed)
  binding    ctypes view of the C ABI in include/ais_b200.h (libais_b200.so, sm_100a; no CPU fallback)
  engine     SearchEngine: stages an index shard into HBM and runs the CUDA path
  query      host-side query parsing / query-vector preparation (the parts that stay Python)
  webui_api  the callables webui.py uses (load_model, find_similar_documents, ...)
  shard      doc-sharded multi-GPU search over torch.distributed
"""
__all__ = ["synth", "binding", "engine", "query", "webui_api", "shard"]
