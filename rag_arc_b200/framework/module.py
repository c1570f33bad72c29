"""``AbstractModule``: what ``AbstractConfig.build()`` returns
(/root/reference framework/module.py:9-11 - a dataclass holding its config)."""
from abc import ABC
from dataclasses import dataclass
from typing import Any


@dataclass
class AbstractModule(ABC):
    config: Any
