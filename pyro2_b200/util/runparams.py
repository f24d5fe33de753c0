"""Runtime parameters: ``[section]`` / ``key = value ; comment`` files and a flat
``section.key`` dictionary, with the call surface of pyro/util/runparams.py:84-232
(load_params, get_param, set_param, command_line_params, print_paramfile, __str__)."""
import os
import re

from . import msg


def _convert(value):
    """int, then float, then stripped string (runparams.py:60-81)"""
    value = value.strip()
    for cast in (int, float):
        try:
            return cast(value)
        except ValueError:
            pass
    return value


class RuntimeParameters:
    _section = re.compile(r"^\[(.*)\]")
    _assign = re.compile(r"^([^=#]+)=([^;]+);{0,1}(.*)")

    def __init__(self):
        self.params = {}
        self.param_comments = {}
        self.used_params = []

    def load_params(self, pfile, *, no_new=False):
        if not os.path.isfile(pfile):
            alt = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), pfile)
            if os.path.isfile(alt):
                pfile = alt
        try:
            with open(pfile) as f:
                lines = f.readlines()
        except OSError:
            msg.fail(f"ERROR: parameter file does not exist: {pfile}")
        section = ""
        for line in lines:
            m = self._section.search(line)
            if m:
                section = m.group(1).strip().lower()
                continue
            m = self._assign.search(line)
            if not m:
                continue
            key = section + "." + m.group(1).strip().lower()
            if no_new and key not in self.params:
                msg.warning(f"warning, key: {key} not defined")
                continue
            self.params[key] = _convert(m.group(2))
            comment = m.group(3).strip()
            if comment == "":
                comment = self.param_comments.get(key, "")
            self.param_comments[key] = comment

    def load_dict(self, table, *, no_new=False):
        """table: {"section.key": value} or {"section.key": (value, comment)}"""
        for key, item in table.items():
            value, comment = item if isinstance(item, tuple) else (item, "")
            if no_new and key not in self.params:
                msg.warning(f"warning, key: {key} not defined")
                continue
            self.params[key] = value
            if comment or key not in self.param_comments:
                self.param_comments[key] = comment

    def command_line_params(self, cmd_strings):
        """``section.key=value`` strings override existing parameters (runparams.py:166-187)"""
        for item in cmd_strings:
            key, value = item.split("=")
            if key not in self.params:
                msg.warning(f"warning, key: {key} not defined")
                continue
            self.params[key] = _convert(value)

    def get_param(self, key):
        if key not in self.used_params:
            self.used_params.append(key)
        if key in self.params:
            return self.params[key]
        raise KeyError(f"ERROR: runtime parameter {key} not found")

    def set_param(self, key, value, *, no_new=True):
        if no_new and key in self.params:
            self.params[key] = value
            return
        if not no_new:
            self.params[key] = value
            self.param_comments.setdefault(key, "")
            return
        raise KeyError(f"ERROR: runtime parameter {key} not found")

    def print_unused_params(self):
        for key in self.params:
            if key not in self.used_params:
                msg.warning(f"parameter {key} never used")

    def print_all_params(self):
        for key in sorted(self.params):
            print(key, "=", self.params[key])
        print(" ")

    def write_params(self, f):
        """write into an open h5py-like group (runparams.py:222-232)"""
        grp = f.create_group("runtime parameters")
        for key in sorted(self.params):
            grp.attrs[key] = self.params[key]

    def __str__(self):
        return "".join(f"{key} = {self.params[key]}\n" for key in sorted(self.params))

    def print_paramfile(self, path="inputs.auto"):
        """dump every parameter in inputs-file syntax (runparams.py:245-275)"""
        try:
            with open(path, "w") as f:
                f.write("# automagically generated parameter file\n")
                section = None
                for key in sorted(self.params):
                    sec, _, name = key.partition(".")
                    if sec != section:
                        section = sec
                        f.write(f"\n[{section}]\n")
                    comment = self.param_comments.get(key, "")
                    f.write(f"{name} = {self.params[key]}" + (f"   ; {comment}" if comment else "") + "\n")
        except OSError:
            pass   # read-only working directory: the dump is a convenience only
