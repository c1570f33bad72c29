"""Boundary value type of the retrieval path.

``Document`` is what every plugin on the path hands around: vector stores keep them in their
docstore, retrievers return lists of them, fusion keys on ``content``.  Field names, order and
defaults follow /root/reference core/utils/data_model.py:4-9 so that objects built by either
side are interchangeable (positional construction included); the helpers below are additions used
by the persistence code of the B200 store.
"""
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, Mapping, Optional


@dataclass
class Document:
    content: str
    metadata: Dict[str, Any] = field(default_factory=dict)
    id: Optional[str] = None

    def to_dict(self) -> Dict[str, Any]:
        """Plain-dict form (JSON-serialisable when the metadata is)."""
        return asdict(self)

    @classmethod
    def from_dict(cls, data: Mapping[str, Any]) -> "Document":
        return cls(content=data["content"], metadata=dict(data.get("metadata") or {}), id=data.get("id"))
