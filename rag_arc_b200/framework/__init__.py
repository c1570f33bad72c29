"""Plugin registry the retrieval plugins sit behind (mirror of /root/reference framework/)."""
from .config import AbstractConfig
from .module import AbstractModule
from .register import Register
from .singleton_decorator import singleton

__all__ = ["AbstractConfig", "AbstractModule", "Register", "singleton"]
