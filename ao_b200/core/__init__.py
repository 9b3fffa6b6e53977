from .config import AOBaseConfig, config_from_dict, config_to_dict  # noqa: F401
