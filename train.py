"""`python train.py --config codes/<name>_config.json` -- same CLI as the reference (train.py:18-74),
running the ELBO hot path on the sm_100a kernels.  Under torchrun (WORLD_SIZE > 1) the run is batch-sharded data
parallel (one process per GPU, NCCL): `batch_size` is the per-rank batch, every rank draws a disjoint shard of each global
batch (same epoch permutation) and its rows of the global noise tensors, batch sums / batch-norm statistics / gradients are
all-reduced, the hyper-prior is fitted on rank 0 and broadcast, and rank 0 alone prints and writes files."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from codes.data_loader import DataGenerator  # noqa: E402
from codes.models import MNISTModel_digit, MNISTModel_fashion, CelebAModel_densenet  # noqa: E402
from codes.trainers import MNISTTrainer_joint_training, CelebATrainer_joint_training  # noqa: E402
from codes.utils import process_config, create_dirs, get_args, save_config  # noqa: E402


def main():
    try:
        args = get_args()
        config = process_config(args.config)
    except Exception:
        print("missing or invalid arguments")
        exit(0)

    if not torch.cuda.is_available():
        raise RuntimeError("train.py needs a CUDA device (sm_100a); there is no CPU path")
    dist_group = None
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    rank = 0
    if int(os.environ.get('WORLD_SIZE', 1)) > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist_group = dist.group.WORLD
        rank = dist.get_rank()
    if rank == 0:
        create_dirs([config['result_dir'], config['checkpoint_dir']])
        save_config(config)
    else:                                   # rank 0 alone prints and writes (the replicas hold identical state)
        sys.stdout = open(os.devnull, 'w')
    if dist_group is not None:
        dist.barrier()

    data = DataGenerator(config, None)
    classes = {'mnist_digit': MNISTModel_digit, 'mnist_fashion': MNISTModel_fashion, 'celeba': CelebAModel_densenet}
    model = classes[config['exp_name']](config, device='cuda:%d' % local_rank, dist_group=dist_group)
    print("Created a VAE model.")
    print("The current dataset is {}, num hidden units: {}.\n".format(config['exp_name'], config['num_hidden_units']))

    if config['TRAIN_VAE'] or config['TRAIN_sigma'] or config['TRAIN_prior']:
        if config['exp_name'] in ('mnist_digit', 'mnist_fashion'):
            trainer_VAE = MNISTTrainer_joint_training(None, model, data, config)
        else:
            trainer_VAE = CelebATrainer_joint_training(None, model, data, config)
        model.load(None, model="VAE")
        if config['prior'] in ("ours", "hierarchical", "vampPrior"):
            model.load(None, model="prior")
        if config['num_epochs'] > 0:
            trainer_VAE.train()
    if dist_group is not None:
        model.engine.release_graphs()      # captured sub-step graphs hold NCCL kernels: drop them before the group goes away
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
