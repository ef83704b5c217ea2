"""Stage driver with the reference's flags and on-disk formats.

Mirrors /root/reference/genewalk/cli.py: same argparse flags (cli.py:49-157), same stages
(node_vectors, null_distribution, statistics; cli.py:194-252), same files in
``<base_folder>/<project>/`` -- ``genes.pkl``, ``multi_graph.pkl``,
``deepwalk_node_vectors_{i}.pkl``, ``deepwalk_node_vectors_rand_{i}.pkl``,
``genewalk_rand_simdists.pkl``, ``genewalk_results.csv`` (``float_format='%.3e'``) -- so the
stages of either implementation can read the other's outputs.

Input assembly (gene-ID mapping, resource download, Pathway Commons / INDRA / SIF + GO assembly,
cli.py:195-197) is OUT OF SCOPE of this build: when the reference package ``genewalk`` and its
dependency goatools are importable, ``read_gene_list`` / ``load_network`` are delegated to it;
otherwise the node_vectors stage expects an assembled network: ``--network_source pickle
--network_file <multi_graph.pkl>`` with ``--genes`` either a ``genes.pkl`` or a text file of ids
(``--id_type custom``).  The ``visual`` stage is not provided.
"""
import argparse
import copy
import gc
import logging
import os
import pickle
import random

import numpy as np

from .replicates import run_walks_many, null_replicates_many, DEFAULT_CONCURRENT
from .perform_statistics import GeneWalk

logger = logging.getLogger('genewalk.cli')

default_base_folder = os.path.join(os.path.expanduser('~/'), 'genewalk')


def create_folder(base_folder, project):
    project_folder = os.path.join(base_folder, project)
    logger.info('Creating project folder at %s' % project_folder)
    if not os.path.exists(project_folder):
        os.makedirs(project_folder)
    return project_folder


def save_pickle(obj, project_folder, prefix):
    fname = os.path.join(project_folder, '%s.pkl' % prefix)
    logger.info('Saving into %s...' % fname)
    with open(fname, 'wb') as fh:
        pickle.dump(obj, fh)


def load_pickle(project_folder, prefix):
    fname = os.path.join(project_folder, '%s.pkl' % prefix)
    logger.info('Loading %s...' % fname)
    with open(fname, 'rb') as fh:
        return pickle.load(fh)


def build_parser():
    parser = argparse.ArgumentParser(
        description='Run GeneWalk on a list of genes provided in a text file (B200 build of the '
                    'representation-learning stages).')
    parser.add_argument('--project', required=True,
                        help='A name for the project which determines the folder within the '
                             'base folder in which the intermediate and final results are written.')
    parser.add_argument('--genes', required=True,
                        help='Path to a text file with a list of genes of interest (or a genes.pkl '
                             'written by a previous run).')
    parser.add_argument('--id_type', default='hgnc_symbol',
                        choices=['hgnc_symbol', 'hgnc_id', 'ensembl_id', 'mgi_id', 'rgd_id',
                                 'entrez_human', 'entrez_mouse', 'custom'])
    parser.add_argument('--stage', default='all',
                        choices=['all', 'node_vectors', 'null_distribution', 'statistics', 'visual'])
    parser.add_argument('--base_folder', default=default_base_folder)
    parser.add_argument('--network_source', default='pc',
                        choices=['pc', 'indra', 'edge_list', 'sif', 'sif_annot', 'sif_full', 'pickle'])
    parser.add_argument('--network_file', default=None)
    parser.add_argument('--nproc', default=1, type=int,
                        help='Accepted for compatibility; the GPU path does not use host workers.')
    parser.add_argument('--nreps_graph', default=3, type=int)
    parser.add_argument('--nreps_null', default=3, type=int)
    parser.add_argument('--alpha_fdr', default=1, type=float)
    parser.add_argument('--dim_rep', default=8, type=int)
    parser.add_argument('--save_dw', default=False, type=bool)
    parser.add_argument('--random_seed', default=None, type=int)
    parser.add_argument('--go_obo', default=None,
                        help='GPU build only: GO OBO file for the GO node names / domains of a '
                             'sif_full network with custom ids (read without the reference package).')
    parser.add_argument('--concurrent', default=DEFAULT_CONCURRENT, type=int,
                        help='GPU build only: replicates in flight at once on one GPU (does not change results).')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    run_main(args)


def _load_inputs(args, project_folder):
    """genes + assembled MultiGraph (cli.py:195-197)."""
    if args.network_source == 'pickle':
        if args.genes.endswith('.pkl'):
            with open(args.genes, 'rb') as fh:
                genes = pickle.load(fh)
        else:
            if args.id_type != 'custom':
                raise ValueError('Without the reference assemblers a plain gene list needs '
                                 '--id_type custom (gene ids must be node ids of the network).')
            with open(args.genes) as fh:
                ids = [ln.strip() for ln in fh if ln.strip()]
            genes = [{'ID': g} for g in dict.fromkeys(ids)]
        with open(args.network_file, 'rb') as fh:
            graph = pickle.load(fh)
        return genes, graph
    if args.network_source == 'sif_full' and args.id_type == 'custom' and not args.genes.endswith('.pkl'):
        # a complete user network with custom ids needs nothing from the reference's resource
        # layer: read it straight into CSR arrays (edgelist.py; nx_mg_assembler.py:368-422)
        from .edgelist import read_network
        with open(args.genes) as fh:
            ids = [ln.strip() for ln in fh if ln.strip()]
        genes = [{'ID': g} for g in dict.fromkeys(ids)]
        return genes, read_network(args.network_file, 'sif_full', genes, obo=args.go_obo)
    try:
        from genewalk.gene_lists import read_gene_list
        from genewalk.nx_mg_assembler import load_network
        from genewalk.resources import ResourceManager
    except ImportError as e:
        raise RuntimeError('network_source %s needs the reference assemblers (package genewalk '
                           'with goatools), which are not importable here (%s); assemble the '
                           'network elsewhere and pass --network_source pickle' %
                           (args.network_source, e))
    rm = ResourceManager(base_folder=args.base_folder)
    genes = read_gene_list(args.genes, args.id_type, rm)
    MG = load_network(args.network_source, args.network_file, genes, resource_manager=rm)
    return genes, MG.graph


def run_main(args):
    project_folder = create_folder(args.base_folder, args.project)
    log_file = os.path.join(project_folder, 'genewalk_%s.log' % args.stage)
    handler = logging.FileHandler(log_file)
    handler.setFormatter(logging.Formatter('%(levelname)s: [%(asctime)s] %(name)s - %(message)s',
                                           datefmt='%Y-%m-%d %H:%M:%S'))
    logging.getLogger().addHandler(handler)
    try:
        # same argument checks, same messages as cli.py:172-185
        if args.network_source in {'indra', 'sif', 'sif_annot', 'sif_full', 'edge_list', 'pickle'}:
            if not args.network_file:
                raise ValueError('The --network_file argument must be provided'
                                 ' when using --network_source %s.' % args.network_source)
        if args.id_type == 'custom':
            if args.network_source not in {'sif_annot', 'sif_full', 'pickle'}:
                raise ValueError('When using --id_type custom, the --network_source'
                                 ' has to be either sif_annot or sif_full, with '
                                 'the --network_file argument pointing to a custom '
                                 'SIF file.')
        if args.random_seed:
            logger.info('Running with random seed %d' % args.random_seed)
            random.seed(a=int(args.random_seed))

        if args.stage in ('all', 'node_vectors'):
            genes, graph = _load_inputs(args, project_folder)
            save_pickle(genes, project_folder, 'genes')
            # multi_graph.pkl always holds an nx.MultiGraph, as the reference writes it
            save_pickle(graph.to_networkx() if hasattr(graph, 'to_networkx') else graph,
                        project_folder, 'multi_graph')
            # cli.py:200-214; the replicates are independent: up to `--concurrent` are in flight at a time
            step = max(1, args.concurrent)
            for i0 in range(0, args.nreps_graph, step):
                cnt = min(step, args.nreps_graph - i0)
                logger.info('%s-%s/%s' % (i0 + 1, i0 + cnt, args.nreps_graph))
                DWs = run_walks_many(graph, cnt, concurrent=step, vector_size=args.dim_rep,
                                     workers=args.nproc)
                for i, DW in enumerate(DWs, start=i0):
                    if args.save_dw:
                        save_pickle(DW, project_folder, 'deepwalk_%d' % (i + 1))
                    nv = copy.deepcopy(DW.model.wv)
                    save_pickle(nv, project_folder, 'deepwalk_node_vectors_%d' % (i + 1))
                del DWs
                gc.collect()

        if args.stage in ('all', 'null_distribution'):
            MG = load_pickle(project_folder, 'multi_graph')
            srd = []
            step = max(1, args.concurrent)
            for i0 in range(0, args.nreps_null, step):   # cli.py:219-236
                cnt = min(step, args.nreps_null - i0)
                logger.info('%s-%s/%s' % (i0 + 1, i0 + cnt, args.nreps_null))
                done = null_replicates_many(MG, cnt, concurrent=step, vector_size=args.dim_rep,
                                            workers=args.nproc)
                for i, (DW, sims) in enumerate(done, start=i0):
                    if args.save_dw:
                        save_pickle(DW, project_folder, 'deepwalk_rand_%d' % (i + 1))
                    if DW.model is not None:
                        save_pickle(copy.deepcopy(DW.model.wv), project_folder,
                                    'deepwalk_node_vectors_rand_%d' % (i + 1))
                    srd.append(sims)
                del done
                gc.collect()
            srd = np.concatenate(srd) if srd else np.zeros(0, np.float32)
            srd = np.sort(np.asarray(srd, dtype=np.float32))   # == np.asarray(sorted(srd)), cli.py:238
            save_pickle(srd, project_folder, 'genewalk_rand_simdists')

        if args.stage in ('all', 'statistics'):
            MG = load_pickle(project_folder, 'multi_graph')
            genes = load_pickle(project_folder, 'genes')
            nvs = [load_pickle(project_folder, 'deepwalk_node_vectors_%d' % (i + 1))
                   for i in range(args.nreps_graph)]
            null_dist = load_pickle(project_folder, 'genewalk_rand_simdists')
            GW = GeneWalk(MG, genes, nvs, null_dist, gene_id_type=args.id_type)
            df = GW.generate_output(alpha_fdr=args.alpha_fdr)
            fname = os.path.join(project_folder, 'genewalk_results.csv')
            logger.info('Saving final results into %s' % fname)
            df.to_csv(fname, index=False, float_format='%.3e')

        if args.stage == 'visual':
            raise NotImplementedError('the visual stage (plots) is outside this build; run the '
                                      'reference\'s --stage visual on genewalk_results.csv')
    finally:
        logging.getLogger().removeHandler(handler)
        handler.close()


if __name__ == '__main__':
    main()
