"""B200-native tensor-network engine behind the qibotn plugin surface."""
__version__ = "0.1.0"

from .backends import MetaBackend  # noqa: E402,F401
