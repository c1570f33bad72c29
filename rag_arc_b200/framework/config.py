"""``AbstractConfig``: pydantic base of every plugin config.

Contract kept from /root/reference framework/config.py:
* each direct subclass must itself declare ``type: Literal["TAG"] = "TAG"`` - checked when the
  class is created (``:25-70``; TypeError otherwise);
* parsed data cannot carry a different tag (``:73-88``; ValueError -> pydantic ValidationError);
* ``build()`` must be overridden (``:17-22``; NotImplementedError).
Nested plugin configs compose through pydantic discriminated unions on ``type``
(framework/config_test.py:37-45, framework/module_test.py:75-109).
"""
from typing import Literal, get_args, get_origin

from pydantic import BaseModel, field_validator


class AbstractConfig(BaseModel):
    def build(self):
        raise NotImplementedError("Subclasses must implement build() method")

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        own = cls.__dict__.get("__annotations__", {})
        name = cls.__name__
        if "type" not in own:
            raise TypeError(f"{name} must declare `type: Literal['TAG'] = 'TAG'`")
        declared = own["type"]
        default = cls.__dict__.get("type")
        if isinstance(declared, str):        # postponed annotations: textual check only
            if not declared.startswith("Literal["):
                raise TypeError(f"{name}.type must be annotated as Literal['TAG']")
            if default is None:
                raise TypeError(f"{name}.type must have a default value")
            return
        if get_origin(declared) is not Literal:
            raise TypeError(f"{name}.type must be annotated as Literal['TAG']")
        tags = get_args(declared)
        if len(tags) != 1 or not isinstance(tags[0], str):
            raise TypeError(f"{name}.type must be Literal['<single string>']")
        if default != tags[0]:
            raise TypeError(f"{name}.type default must equal {tags[0]!r}")

    @field_validator("type", check_fields=False)
    @classmethod
    def _tag_must_match(cls, value):
        expected = cls.__dict__.get("type")
        if "type" in cls.__annotations__ and expected is not None and value != expected:
            raise ValueError(f"type must be {expected!r}")
        return value
