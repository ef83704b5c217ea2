"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/genewalk_b200.h declares; the ctypes table mirrors the header; no compute is called."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'genewalk_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    decl = re.findall(r'\b(?:int|int64_t|long long|const char \*)\s*(gw_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S)
    return {name: [a.strip() for a in args.split(',')] if args.strip() != 'void' else []
            for name, args in decl}


def test_header_declares_the_hot_path():
    fns = declared_functions()
    for name in ('gw_walk_uniform', 'gw_sgns_train', 'gw_alias_build', 'gw_config_model',
                 'gw_csr_from_edges', 'gw_cosine_pairs', 'gw_null_similarities', 'gw_sort_f32',
                 'gw_psim', 'gw_bh_segmented', 'gw_log_stats', 'gw_count_tokens', 'gw_vocab_rank',
                 'gw_sgns_init', 'gw_mean_sem_f32', 'gw_walk_count', 'gw_sgns_plan',
                 'gw_sgns_plan_capacity', 'gw_sgns_auto_shape'):
        assert name in fns
    # every entry point cites the reference call site it replaces
    text = open(HEADER).read()
    for cite in ('deepwalk.py:50-96', 'deepwalk.py:135-139', 'null_distributions.py:23-25',
                 'null_distributions.py:33-48', 'perform_statistics.py:102',
                 'perform_statistics.py:247-258', 'perform_statistics.py:112', 'cli.py:238'):
        assert cite in text


def test_library_loads_and_exports_every_declared_symbol():
    from genewalk_b200 import _lib
    lib = _lib.load()
    assert os.path.exists(_lib.library_path())
    fns = declared_functions()
    for name, args in fns.items():
        assert hasattr(lib, name), 'missing export %s' % name
    assert lib.gw_version() == _lib.ABI_VERSION == 200
    assert isinstance(_lib.last_error(), str)
    assert _lib.launch_count() >= 0


def test_ctypes_table_matches_header_arity():
    from genewalk_b200 import _lib
    fns = declared_functions()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in fns, '%s bound but not declared in the header' % name
        assert len(argtypes) == len(fns[name]), '%s: %d ctypes args vs %d declared' % (
            name, len(argtypes), len(fns[name]))
    for name, argtypes in _lib.INT64_RETURNS.items():
        assert len(argtypes) == len(fns[name])
    bound = set(_lib.SIGNATURES) | set(_lib.INT64_RETURNS) | set(_lib.INT_RETURNS) | \
        {'gw_last_error', 'gw_launch_count'}
    assert set(fns) == bound


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import networkx as nx
    from genewalk_b200 import DeepWalk, _lib
    mg = nx.MultiGraph()
    mg.add_edge('a', 'b')
    with pytest.raises(_lib.GeneWalkCudaError):
        DeepWalk(mg).get_walks()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'genewalk_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'libgw_oracle' not in text and 'oracle.lib' not in text, f
