"""Config / directory helpers with the reference's interface (codes/utils.py:11-124).

Same JSON keys, same derived `summary_dir / result_dir / checkpoint_dir`, same save-name
template, same `-c/--config` flag.  Optional NEW keys (all default to the reference behaviour):
`synthetic` (bool), `synthetic_n_train`, `synthetic_n_val`, `data_path` (MNIST idx/npz folder),
`cuda_graphs` (bool), `seed`.
"""
import argparse
import json
import os
from datetime import datetime


def get_config_from_json(json_file):
    with open(json_file, 'r') as f:
        return json.load(f)


def save_config(config):
    stamp = datetime.now().strftime("%d-%b-%Y-%H-%M")
    filename = os.path.join(config['checkpoint_dir'], 'training_config_{}.txt'.format(stamp))
    with open(filename, 'w') as f:
        f.write(json.dumps(config))
    print('The current config is saved at {}'.format(filename))


def experiment_name(config):
    """prior-{prior}-{H}-{C}-{R}-{act}-{layers}-mixture-{K}  (codes/utils.py:49-56)"""
    keys = ('prior', 'num_hidden_units', 'code_size', 'representation_size', 'inner_activation',
            'n_layers_inner_VAE', 'n_mixtures')
    return 'prior-{}-{}-{}-{}-{}-{}-mixture-{}'.format(*[config[k] for k in keys])


def process_config(json_file):
    config = get_config_from_json(json_file)
    print("The current config is:\n{}\n".format(config))
    save_name = experiment_name(config)
    print("Experiment results will be saved at:\n{}\n".format(save_name))
    if config['load_dir'] == "default":
        save_dir = "./experiments/{}/batch-{}".format(config['exp_name'], config['batch_size'])
        for key, leaf in (('summary_dir', 'summary/'), ('result_dir', 'result/'), ('checkpoint_dir', 'checkpoint/')):
            config[key] = os.path.join(save_dir, save_name, leaf)
    else:
        config['summary_dir'] = "./figures/{}/summary/".format(config['exp_name'])
        config['result_dir'] = "./figures/{}/result/".format(config['exp_name'])
        config['checkpoint_dir'] = os.path.join(config['load_dir'], config['exp_name'])
    print("Models will be saved / loaded at:\n{}".format(config['checkpoint_dir']))
    print("Results will be saved at:\n{}\n".format(config['result_dir']))
    return config


def create_dirs(dirs):
    try:
        for d in dirs:
            os.makedirs(d, exist_ok=True)
        return 0
    except Exception as err:  # same contract as the reference: report and exit(-1)
        print("Creating directories error: {0}".format(err))
        exit(-1)


def count_trainable_variables(model, scope_name):
    """Number of trainable parameters under a variable scope of `model` (a host.models class)."""
    total = 0
    for name, t in model.engine.named_parameters():
        if name.split('/')[0] == scope_name:
            total += t.numel()
    print('The total number of trainable parameters in the {} model is: {}k.'.format(
        scope_name, round(total / 1000, 2)))
    return total


def get_args(argv=None):
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument('-c', '--config', metavar='C', default='None', help='The Configuration file')
    return p.parse_args(argv)
