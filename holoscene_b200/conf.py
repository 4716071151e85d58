"""Minimal HOCON-subset reader with the pyhocon ConfigTree accessors the reference uses
(training/holoscene_train.py:48,104-169; model/network.py:757-770): nested `name { ... }` /
`name = { ... }` blocks, `key = value`, lists, `#` and `//` comments, unquoted strings, dotted
look-ups, get_int / get_float / get_bool / get_string / get_list / get_config with defaults.
pyhocon itself is used when it is installed; this keeps the confs/*.conf format usable without it.
"""
from __future__ import annotations

import re

_MISSING = object()


class ConfigTree(dict):
    def _walk(self, key):
        node = self
        for part in key.split("."):
            if not isinstance(node, dict) or part not in node:
                raise KeyError(key)
            node = node[part]
        return node

    def get(self, key, default=_MISSING):
        try:
            return self._walk(key)
        except KeyError:
            if default is _MISSING:
                raise
            return default

    def _typed(self, key, default, cast):
        try:
            return cast(self._walk(key))
        except KeyError:
            if default is _MISSING:
                raise
            return default

    def get_int(self, key, default=_MISSING):
        return self._typed(key, default, int)

    def get_float(self, key, default=_MISSING):
        return self._typed(key, default, float)

    def get_bool(self, key, default=_MISSING):
        def cast(v):
            if isinstance(v, str):
                return v.strip().lower() in ("true", "yes", "on", "1")
            return bool(v)
        return self._typed(key, default, cast)

    def get_string(self, key, default=_MISSING):
        return self._typed(key, default, str)

    def get_list(self, key, default=_MISSING):
        return self._typed(key, default, list)

    def get_config(self, key, default=_MISSING):
        v = self._typed(key, default, lambda x: x)
        return v

    def put(self, key, value):
        node = self
        parts = key.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, ConfigTree())
        node[parts[-1]] = value


_TOKEN = re.compile(r"""\s*(?:(?P<brace>[{}\[\],=:])|"(?P<q>(?:[^"\\]|\\.)*)"|(?P<w>[^\s{}\[\],=:"]+))""")


def _scalar(tok: str):
    low = tok.lower()
    if low in ("true", "yes", "on"):
        return True
    if low in ("false", "no", "off"):
        return False
    if low == "null":
        return None
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


def _tokenize(text: str):
    out = []
    for line in text.splitlines():
        # strip comments outside quotes
        buf, inq = [], False
        i = 0
        while i < len(line):
            ch = line[i]
            if ch == '"':
                inq = not inq
            if not inq and (ch == "#" or line.startswith("//", i)):
                break
            buf.append(ch)
            i += 1
        s = "".join(buf)
        pos = 0
        while pos < len(s):
            m = _TOKEN.match(s, pos)
            if not m:
                break
            pos = m.end()
            if m.group("brace"):
                out.append(("p", m.group("brace")))
            elif m.group("q") is not None:
                out.append(("s", m.group("q")))
            else:
                out.append(("w", m.group("w")))
        out.append(("nl", "\n"))
    return out


class _Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def skip_nl(self):
        while self.peek()[0] == "nl" or self.peek() == ("p", ","):
            self.i += 1

    def parse_object(self, closing):
        tree = ConfigTree()
        while True:
            self.skip_nl()
            kind, val = self.peek()
            if kind == "eof":
                if closing:
                    raise ValueError("unterminated { in conf")
                return tree
            if (kind, val) == ("p", "}"):
                self.next()
                return tree
            key = self.next()[1]
            kind, val = self.peek()
            while kind == "nl":      # `name` newline `{`
                self.next()
                kind, val = self.peek()
            if (kind, val) in (("p", "="), ("p", ":")):
                self.next()
                value = self.parse_value()
            elif (kind, val) == ("p", "{"):
                value = self.parse_value()
            else:
                raise ValueError(f"conf: expected '=' or '{{' after key {key!r}")
            if isinstance(value, ConfigTree) and isinstance(tree.get(key, None), ConfigTree):
                tree[key].update(value)
            else:
                tree.put(key, value)

    def parse_value(self):
        kind, val = self.next()
        while kind == "nl":
            kind, val = self.next()
        if (kind, val) == ("p", "{"):
            return self.parse_object(True)
        if (kind, val) == ("p", "["):
            items = []
            while True:
                self.skip_nl()
                if self.peek() == ("p", "]"):
                    self.next()
                    return items
                items.append(self.parse_value())
        if kind == "s":
            return val
        if kind == "w":
            # unquoted strings may contain spaces up to end of line
            words = [val]
            while self.peek()[0] == "w":
                words.append(self.next()[1])
            return _scalar(words[0]) if len(words) == 1 else " ".join(words)
        raise ValueError(f"conf: unexpected token {val!r}")


def parse_string(text: str) -> ConfigTree:
    return _Parser(_tokenize(text)).parse_object(False)


def parse_file(path: str) -> ConfigTree:
    try:
        from pyhocon import ConfigFactory  # type: ignore
        return ConfigFactory.parse_file(path)
    except ImportError:
        with open(path) as f:
            return parse_string(f.read())


def from_dict(d) -> ConfigTree:
    t = ConfigTree()
    for k, v in d.items():
        t[k] = from_dict(v) if isinstance(v, dict) else v
    return t
