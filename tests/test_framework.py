"""CPU: the registry / config plugin API behaves like the reference's framework/ (the cases mirror
/root/reference framework/register_test.py, config_test.py and module_test.py)."""
import json
import os
import tempfile
from dataclasses import dataclass
from typing import Annotated, List, Literal, Union
from unittest.mock import patch

import pytest
from pydantic import Field, ValidationError

from rag_arc_b200.framework import AbstractConfig, AbstractModule, Register


class LeafA(AbstractConfig):
    type: Literal["A"] = "A"
    x: int

    def build(self):
        return LeafModule(config=self)


class LeafB(AbstractConfig):
    type: Literal["B"] = "B"
    y: str = "b"

    def build(self):
        return LeafModule(config=self)


@dataclass
class LeafModule(AbstractModule):
    def describe(self):
        return self.config.type


class Parent(AbstractConfig):
    type: Literal["P"] = "P"
    child: Annotated[Union[LeafA, LeafB], Field(discriminator="type")]
    many: List[Annotated[Union[LeafA, LeafB], Field(discriminator="type")]] = []

    def build(self):
        return ParentModule(config=self)


@dataclass
class ParentModule(AbstractModule):
    def children(self):
        return [self.config.child.build()] + [c.build() for c in self.config.many]


def _tmp(obj):
    f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    f.write(obj if isinstance(obj, str) else json.dumps(obj))
    f.close()
    return f.name


@pytest.fixture
def reg():
    r = Register()
    r.registrations.clear()
    return r


def test_register_valid_and_get_object(reg):
    path = _tmp({"type": "A", "x": 3})
    reg.register(path, "leaf", LeafA)
    assert reg.get_object("leaf").config.x == 3
    assert Register() is reg                      # singleton
    os.unlink(path)


def test_missing_file_propagates_and_unknown_name_is_keyerror(reg):
    with pytest.raises(FileNotFoundError):
        reg.register("/no/such/file.json", "x", LeafA)
    with pytest.raises(KeyError):
        reg.get_object("nope")


@pytest.mark.parametrize("content", ["", "{not json", json.dumps({"type": "A"}), json.dumps({"type": "B", "x": 1}),
                                     json.dumps({"type": "A", "x": "not-an-int"})])
def test_bad_content_is_caught_printed_and_not_registered(reg, content):
    path = _tmp(content)
    with patch("builtins.print") as pr:
        reg.register(path, "bad", LeafA)
    assert "bad" not in reg.registrations
    assert pr.called and "Error registering bad" in pr.call_args[0][0]
    os.unlink(path)


def test_same_name_overwrites_and_multiple_types(reg):
    p1, p2 = _tmp({"type": "A", "x": 1}), _tmp({"type": "B", "y": "z"})
    reg.register(p1, "app", LeafA)
    reg.register(p2, "app", LeafB)
    assert reg.get_object("app").describe() == "B"
    reg.register(p1, "other", LeafA)
    assert set(reg.registrations) == {"app", "other"}


def test_discriminated_union_and_nested_build(reg):
    cfg = Parent(child={"type": "B"}, many=[{"type": "A", "x": 1}, {"type": "B", "y": "q"}])
    assert isinstance(cfg.child, LeafB)
    kinds = [m.describe() for m in cfg.build().children()]
    assert kinds == ["B", "A", "B"]
    with pytest.raises(ValidationError):
        Parent(child={"type": "C"})


def test_tag_rules_enforced_at_class_creation_and_at_parse_time():
    with pytest.raises(TypeError):
        class NoTag(AbstractConfig):
            x: int = 0
    with pytest.raises(TypeError):
        class NotLiteral(AbstractConfig):
            type: str = "T"
    with pytest.raises(TypeError):
        class WrongDefault(AbstractConfig):
            type: Literal["T"] = "U"
    with pytest.raises(ValidationError):
        LeafA(type="B", x=1)
    with pytest.raises(NotImplementedError):
        class NoBuild(AbstractConfig):
            type: Literal["N"] = "N"
        NoBuild().build()


def test_b200_configs_parse_and_nest():
    from rag_arc_b200.configs import HybridRetrieverConfig
    cfg = HybridRetrieverConfig(**{
        "type": "b200_hybrid_retriever",
        "retrievers": [
            {"type": "b200_dense_retriever",
             "vectorstore": {"type": "b200_vector_store", "embedding": {"type": "hash_embeddings", "dim": 64},
                             "dtype": "bfloat16", "corpus_path": "corpus.jsonl"}},
            {"type": "b200_bm25_retriever", "corpus_path": "corpus.jsonl", "k": 7}],
        "fusion": {"type": "rrf", "k": 60.0}})
    assert cfg.retrievers[0].vectorstore.embedding.dim == 64 and cfg.retrievers[1].k == 7
    with pytest.raises(ValidationError):
        HybridRetrieverConfig(retrievers=[{"type": "nope"}])
