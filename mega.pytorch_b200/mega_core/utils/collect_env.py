"""utils/collect_env.py:7-14: torch's environment report plus the B200 library's identity"""
from torch.utils.collect_env import get_pretty_env_info


def collect_env_info():
    from .. import _lib
    return get_pretty_env_info() + "\n        libmega_b200: %s (ABI v%d)" % (_lib.LIB_PATH, _lib.lib.mega_abi_version())
