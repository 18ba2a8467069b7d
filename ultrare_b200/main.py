"""B200 mirror of the reference's main.py: identical argparse surface and dispatch
(main.py:5-14,17-70), plus optional flags whose defaults preserve the reference behaviour.
"""
import argparse

parser = argparse.ArgumentParser()
parser.add_argument('--dataset', type=str, default='ml1m', help='dataset name')
parser.add_argument('--epoch', type=int, default=50, help='number of epochs')
parser.add_argument('--worker', type=int, default=24, help='number of CPU workers')
parser.add_argument('--verbose', type=int, default=1, help='verbose type')
parser.add_argument('--group', type=int, default=2, help='number of groups')
parser.add_argument('--layer', nargs='+', type=int, default=[64, 32], help='setting of layers')   # Appendix A12
parser.add_argument('--learn', type=str, default='sisa', help='type of learning and unlearning')
parser.add_argument('--delper', type=int, default=2, help='deleted user proportion')
parser.add_argument('--deltype', type=str, default='rand', help='deletion type')
# additions (not in the reference)
parser.add_argument('--synth', action='store_true',
                    help='write a deterministic synthetic ML-1M-shaped data/ml1m/squ0_{train,test}.csv if absent')
parser.add_argument('--epoch-eval', type=str, default=None, choices=['faithful', 'final', 'none'],
                    help='per-epoch in-training evaluation fidelity of the SISA path')


def main(argv=None):
    args = parser.parse_args(argv)
    assert args.dataset in ['ml1m', 'toy']
    dataset = args.dataset
    assert args.epoch > 0
    epochs = args.epoch
    assert args.worker > 0
    n_worker = args.worker
    assert args.verbose in [0, 1, 2]
    verbose = args.verbose
    assert args.group >= 0
    n_group = args.group
    for i in args.layer:
        assert type(i) == int
    layers = args.layer
    assert args.learn in ['sisa']
    learn_type = args.learn
    assert args.delper in [2, 5]
    del_per = args.delper
    assert args.deltype in ['rand']
    del_type = args.deltype

    import os
    if args.epoch_eval:
        os.environ['ULTRARE_EPOCH_EVAL'] = args.epoch_eval
    from . import dist as udist
    udist.init_from_env()
    from .config import InsParam, Instance
    if args.synth:
        from . import synth
        synth.ensure_dataset(dataset)

    param = InsParam(dataset, epochs, n_worker, layers, n_group, del_per, del_type)
    ins = Instance(param)
    if n_group == 0:
        ins.runFull(is_save=True, verbose=verbose)
    else:
        for group_type in ['emb-ot']:
            ins.runGroup(is_save=True, learn_type=learn_type, group_type=group_type, n_group=n_group, verbose=verbose)
    return ins


if __name__ == '__main__':
    main()
