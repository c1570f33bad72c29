"""``Register``: name -> built module, loaded from JSON config files.

Same behaviour as /root/reference framework/register.py:15-26: ``open()`` errors propagate
(FileNotFoundError), anything wrong *inside* the file (JSON, validation, build) is caught and
printed and nothing is registered; registering an existing name overwrites; ``get_object`` of
an unknown name raises KeyError.
"""
import json

from .singleton_decorator import singleton


@singleton
class Register:
    def __init__(self):
        self.registrations = {}

    def register(self, config_path, app_name, config_type):
        with open(config_path, "r") as fh:
            try:
                cfg = config_type(**json.loads(fh.read()))
                self.registrations[app_name] = cfg.build()
            except Exception as exc:  # noqa: BLE001 - the reference swallows every error here
                print(f"Error registering {app_name}, the config file is not valid\n {exc}")

    def get_object(self, app_name):
        return self.registrations[app_name]
