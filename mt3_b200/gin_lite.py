"""A reader for the subset of gin-config syntax the MT3 model configs use.

gin-config is not installed here; the reference's gin/model.gin, mt3.gin and ismir2021.gin
(or this package's copies of their inference-relevant bindings) stay the source of truth
for hyper-parameters.  Supported: comments, `import`/`from ... import` lines (ignored),
`include` (resolved relative to the file), macros `NAME = value`, bindings
`scope/mod.fn.param = value`, block bindings `mod.Cls:` + indented `param = value`,
values = Python literals, `%MACRO` and `@callable` / `@callable()` references.
"""
from __future__ import annotations

import ast
import os
import re
from typing import Any, Dict, Iterable


class Macro:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"%{self.name}"


class Ref:
    def __init__(self, name, call):
        self.name, self.call = name, call

    def __repr__(self):
        return f"@{self.name}{'()' if self.call else ''}"


_TOKEN = re.compile(r"(?<![\w'\"])([%@])([A-Za-z_][\w./]*)(\(\))?")


def _parse_value(text: str) -> Any:
    text = text.strip()
    refs: Dict[str, Any] = {}

    def sub(m):
        key = f"__gin_ref_{len(refs)}__"
        refs[key] = Macro(m.group(2)) if m.group(1) == '%' else Ref(m.group(2), bool(m.group(3)))
        return repr(key)

    py = _TOKEN.sub(sub, text)
    val = ast.literal_eval(py)

    def restore(v):
        if isinstance(v, str) and v in refs:
            return refs[v]
        if isinstance(v, (list, tuple)):
            return type(v)(restore(x) for x in v)
        if isinstance(v, dict):
            return {restore(k): restore(x) for k, x in v.items()}
        return v

    return restore(val)


class Config:
    """Parsed bindings: `macros[name]`, `bindings['mod.Cls'][param]`."""

    def __init__(self):
        self.macros: Dict[str, Any] = {}
        self.bindings: Dict[str, Dict[str, Any]] = {}

    def parse_file(self, path: str) -> "Config":
        with open(path) as f:
            self.parse_lines(f.read().splitlines(), os.path.dirname(path))
        return self

    def parse_lines(self, lines: Iterable[str], base_dir: str = ".") -> "Config":
        block = None
        pending = ""
        for raw in lines:
            line = raw.split('#', 1)[0].rstrip() if "'" not in raw and '"' not in raw else _strip_comment(raw)
            if not line.strip():
                continue
            if pending:
                line = pending + " " + line.strip()
                pending = ""
            if _unbalanced(line):
                pending = line
                continue
            indented = line[0] in " \t"
            s = line.strip()
            if s.startswith("import ") or s.startswith("from "):
                continue
            if s.startswith("include "):
                inc = ast.literal_eval(s[len("include "):].strip())
                p = inc if os.path.isabs(inc) else os.path.join(base_dir, inc)
                if not os.path.exists(p):   # the reference writes include paths relative to its repo root
                    p = os.path.join(base_dir, os.path.basename(inc))
                self.parse_file(p)
                continue
            if s.endswith(":") and "=" not in s:
                block = s[:-1].strip()
                continue
            if "=" not in s:
                raise ValueError(f"gin_lite: cannot parse line {raw!r}")
            lhs, rhs = s.split("=", 1)
            lhs = lhs.strip()
            if indented and block:
                self.bindings.setdefault(_strip_scope(block), {})[lhs] = _parse_value(rhs)
                continue
            block = None
            if "." in lhs:
                target, param = lhs.rsplit(".", 1)
                self.bindings.setdefault(_strip_scope(target), {})[param] = _parse_value(rhs)
            else:
                self.macros[lhs] = _parse_value(rhs)
        return self

    def parse_bindings(self, bindings: Iterable[str]) -> "Config":
        return self.parse_lines(list(bindings))

    def macro(self, name: str, default=None, _depth=0):
        v = self.macros.get(name, default)
        while isinstance(v, Macro) and _depth < 16:
            v = self.macros.get(v.name, default)
            _depth += 1
        return v

    def binding(self, target: str, param: str, default=None):
        """Looks `target.param` up by suffix match on the configurable name, so
        'network.T5Config' also finds 'mt3.network.T5Config'."""
        for k, d in self.bindings.items():
            if k == target or k.endswith("." + target) or target.endswith("." + k):
                if param in d:
                    v = d[param]
                    return self.macro(v.name, default) if isinstance(v, Macro) else v
        return default

    def params(self, target: str) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for k, d in self.bindings.items():
            if k == target or k.endswith("." + target) or target.endswith("." + k):
                for p, v in d.items():
                    out[p] = self.macro(v.name) if isinstance(v, Macro) else v
        return out


def _strip_scope(name: str) -> str:
    return name.split("/")[-1].strip()


def _strip_comment(raw: str) -> str:
    out, q = [], None
    for ch in raw:
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == '#':
            break
        out.append(ch)
    return "".join(out).rstrip()


def _unbalanced(s: str) -> bool:
    depth, q = 0, None
    for ch in s:
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
    return depth > 0


def parse_config_files_and_bindings(config_files, bindings=()) -> Config:
    cfg = Config()
    for f in config_files:
        cfg.parse_file(f)
    cfg.parse_bindings(bindings)
    return cfg
