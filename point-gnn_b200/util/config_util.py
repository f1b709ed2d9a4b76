"""Configuration files - mirror of the reference's ``util/config_util.py`` (JSON load / save)."""
import json


def load_config(filename):
    """config_util.py:5-9."""
    with open(filename, 'r') as f:
        return json.load(f)


def save_config(filename, config):
    """config_util.py:11-14."""
    with open(filename, 'w') as f:
        json.dump(config, f, sort_keys=True, indent=4)


load_train_config = load_config
save_train_config = save_config
