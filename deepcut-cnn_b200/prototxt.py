"""Protobuf text-format reader used by the Python-side harness (weight synthesis,
caffemodel writer, bench).  The product's real loader is the C++ one in
``caffe_host/`` (src/proto_text.cpp); this mirrors its behaviour for harness code.

A message is ``Msg`` (dict subclass): field name -> list of values in file order;
``m.one(name, default)`` gives protobuf "optional" semantics (last wins).
Replaces: ReadProtoFromTextFile, reference src/caffe/util/io.cpp:34-42.
"""


class Msg(dict):
    def one(self, name, default=None):
        v = self.get(name)
        return v[-1] if v else default

    def rep(self, name):
        return self.get(name, [])

    def sub(self, name):
        v = self.get(name)
        return v[-1] if v else Msg()


class _Lexer:
    PUNCT = "{}<>[]:,;"

    def __init__(self, text):
        self.t = text
        self.i = 0
        self.n = len(text)

    def next(self):
        t, n = self.t, self.n
        while self.i < n:
            c = t[self.i]
            if c == "#":
                while self.i < n and t[self.i] != "\n":
                    self.i += 1
            elif c.isspace():
                self.i += 1
            else:
                break
        if self.i >= n:
            return None
        c = t[self.i]
        if c in self.PUNCT:
            self.i += 1
            return ("p", c)
        if c in "\"'":
            j = self.i + 1
            buf = []
            while j < n and t[j] != c:
                if t[j] == "\\" and j + 1 < n:
                    j += 1
                    buf.append({"n": "\n", "t": "\t"}.get(t[j], t[j]))
                else:
                    buf.append(t[j])
                j += 1
            if j >= n:
                raise ValueError("prototxt: unterminated string at %d" % self.i)
            self.i = j + 1
            return ("s", "".join(buf))
        j = self.i
        while j < n and not t[j].isspace() and t[j] not in self.PUNCT and t[j] not in "\"'#":
            j += 1
        tok = t[self.i:j]
        self.i = j
        return ("w", tok)


def _value(kind, tok):
    if kind == "s":
        return tok
    for conv in (lambda s: int(s, 0), float):
        try:
            return conv(tok)
        except ValueError:
            pass
    return {"true": True, "false": False}.get(tok, tok)


def _message(lx, closer):
    m = Msg()
    while True:
        tk = lx.next()
        if tk is None:
            if closer:
                raise ValueError("prototxt: missing '%s'" % closer)
            return m
        kind, tok = tk
        if kind == "p":
            if tok == closer:
                return m
            if tok in ",;":
                continue
            raise ValueError("prototxt: unexpected '%s'" % tok)
        name = tok
        tk = lx.next()
        if tk == ("p", ":"):
            tk = lx.next()
        if tk is None:
            raise ValueError("prototxt: field '%s' has no value" % name)
        kind, tok = tk
        if kind == "p" and tok in "{<":
            m.setdefault(name, []).append(_message(lx, "}" if tok == "{" else ">"))
        elif kind == "p" and tok == "[":
            while True:
                tk = lx.next()
                if tk is None:
                    raise ValueError("prototxt: missing ']'")
                if tk == ("p", "]"):
                    break
                if tk == ("p", ","):
                    continue
                m.setdefault(name, []).append(_value(*tk))
        elif kind == "p":
            raise ValueError("prototxt: bad value for '%s'" % name)
        else:
            m.setdefault(name, []).append(_value(kind, tok))


def parse(text):
    return _message(_Lexer(text), None)


def parse_file(path):
    with open(path) as f:
        return parse(f.read())
