"""Configurator: ini-style configuration + ``--key=value`` command line, typed on access.

Mirror of the reference's util/configurator.py (same constructor, lookup order, typing rules,
``params_str`` and ``__str__`` output) so that ``main.py`` driven by ``NeuRec.properties`` and
``conf/<recommender>.properties`` behaves identically.
"""
import os
import sys
from collections import OrderedDict
from configparser import ConfigParser

_SPECIAL = set('/\\":*?<>|\t')
_TYPES = (str, int, float, list, tuple, bool, type(None))


class Configurator(object):
    def __init__(self, config_file, default_section="default"):
        # util/configurator.py:43-67
        if not os.path.isfile(config_file):
            raise FileNotFoundError("There is not config file named '%s'!" % config_file)
        self._default_section = default_section
        self.cmd_arg = self._read_cmd_arg()
        self.lib_arg = self._read_config_file(config_file)
        arg_file = os.path.join(self.lib_arg["config_dir"], self.lib_arg["recommender"] + ".properties")
        self.alg_arg = self._read_config_file(arg_file)

    @staticmethod
    def _read_cmd_arg():
        # util/configurator.py:69-78
        cmd = OrderedDict()
        if "ipykernel_launcher" in sys.argv[0]:
            return cmd
        for arg in sys.argv[1:]:
            if not arg.startswith("--"):
                raise SyntaxError("Commend arg must start with '--', but '%s' is not!" % arg)
            name, value = arg[2:].split("=")
            cmd[name] = value
        return cmd

    def _read_config_file(self, filename):
        # util/configurator.py:80-101: single section is used whatever its name; command-line
        # values override keys that are present in the file.
        parser = ConfigParser()
        parser.optionxform = str
        parser.read(filename, encoding="utf-8")
        sections = parser.sections()
        if not sections:
            raise ValueError("'%s' is empty!" % filename)
        if len(sections) == 1:
            section = sections[0]
        elif self._default_section in sections:
            section = self._default_section
        else:
            raise ValueError("'%s' has more than one sections but there is no section named '%s'"
                             % (filename, self._default_section))
        args = OrderedDict(parser[section].items())
        for key in self.cmd_arg:
            if key in args:
                args[key] = self.cmd_arg[key]
        return args

    def params_str(self):
        # util/configurator.py:103-114
        body = "_".join("{}={}".format(k, v) for k, v in self.alg_arg.items() if len(v) < 20)
        body = "".join(c if c not in _SPECIAL else "_" for c in body)
        return "%s_%s" % (self["recommender"], body)

    def __getitem__(self, item):
        # util/configurator.py:116-142: lib -> alg -> cmd, then eval() typing
        if not isinstance(item, str):
            raise TypeError("index must be a str")
        for table in (self.lib_arg, self.alg_arg, self.cmd_arg):
            if item in table:
                param = table[item]
                break
        else:
            raise KeyError("There are not the parameter named '%s'" % item)
        try:
            value = eval(param)
            if not isinstance(value, _TYPES):
                value = param
        except Exception:
            low = param.lower()
            value = True if low == "true" else False if low == "false" else param
        return value

    def __getattr__(self, item):
        return self[item]

    def __contains__(self, o):
        return o in self.lib_arg or o in self.alg_arg or o in self.cmd_arg

    def __str__(self):
        lib = "\n".join("{}={}".format(k, v) for k, v in self.lib_arg.items())
        alg = "\n".join("{}={}".format(k, v) for k, v in self.alg_arg.items())
        return "\n\nNeuRec hyperparameters:\n%s\n\n%s's hyperparameters:\n%s\n" % (lib, self["recommender"], alg)

    __repr__ = __str__
