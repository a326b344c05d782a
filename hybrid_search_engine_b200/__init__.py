"""hybrid_search_engine_b200 -- B200-native hybrid scoring path, drop-in for the pipelines API of
coff33ninja/hybrid-search-engine (``create_pipeline(...).index() / .search()``).

Importing the package does not touch the GPU; the CUDA library is loaded on first use and there is
no CPU fallback.
"""
__version__ = "0.1.0"

_LAZY = {
    "create_pipeline": "pipelines", "PipelineResult": "pipelines", "BasePipeline": "pipelines",
    "BasicPipeline": "pipelines", "BM25Pipeline": "pipelines", "HybridBM25Pipeline": "pipelines",
    "MultiStagePipeline": "pipelines", "DiversityPipeline": "pipelines",
    "Searcher": "core", "BM25": "bm25", "BM25Okapi": "bm25", "BM25Plus": "bm25",
    "extract_tokens": "extractor", "preprocess_text": "extractor", "STOPWORDS": "extractor",
    "DeviceIndex": "index", "SearchEngine": "engine", "QueryBatch": "engine",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
