"""Minimal protobuf text-format reader for the oracle (test infrastructure).

Restates what ``ReadProtoFromTextFile`` (reference src/caffe/util/io.cpp:34-42,
i.e. google::protobuf::TextFormat::Parse) yields for a V2 ``layer {}`` net:
every message becomes ``dict[str, list]`` (all fields kept as repeated lists,
scalars converted to int/float/str/bool-ish identifiers).
"""
import re

_TOKEN = re.compile(r"""
    \s+ | \#[^\n]* |                      # whitespace / comments
    (?P<str>"(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*') |
    (?P<punct>[{}:<>,;\[\]]) |
    (?P<word>[^\s{}:<>,;\[\]"'#]+)
""", re.X)


def _tokens(text):
    pos = 0
    out = []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError("prototxt: bad character at offset %d" % pos)
        pos = m.end()
        if m.lastgroup == "str":
            out.append(("str", re.sub(r"\\(.)", lambda e: {"n": "\n", "t": "\t"}.get(e.group(1), e.group(1)),
                                      m.group("str")[1:-1])))
        elif m.lastgroup == "punct":
            out.append(("p", m.group("punct")))
        elif m.lastgroup == "word":
            out.append(("w", m.group("word")))
    return out


def _scalar(kind, tok):
    if kind == "str":
        return tok
    try:
        return int(tok, 0)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        pass
    if tok == "true":
        return True
    if tok == "false":
        return False
    return tok  # enum identifier


def _parse_msg(toks, i, closer):
    msg = {}
    while i < len(toks):
        kind, tok = toks[i]
        if kind == "p" and tok == closer:
            return msg, i + 1
        if kind == "p" and tok in ",;":
            i += 1
            continue
        if kind != "w":
            raise ValueError("prototxt: expected field name, got %r" % (tok,))
        name = tok
        i += 1
        kind, tok = toks[i]
        if kind == "p" and tok == ":":
            i += 1
            kind, tok = toks[i]
        if kind == "p" and tok in "{<":
            sub, i = _parse_msg(toks, i + 1, "}" if tok == "{" else ">")
            msg.setdefault(name, []).append(sub)
        elif kind == "p" and tok == "[":
            i += 1
            while not (toks[i][0] == "p" and toks[i][1] == "]"):
                if not (toks[i][0] == "p" and toks[i][1] == ","):
                    msg.setdefault(name, []).append(_scalar(*toks[i]))
                i += 1
            i += 1
        else:
            msg.setdefault(name, []).append(_scalar(kind, tok))
            i += 1
    if closer is not None:
        raise ValueError("prototxt: unterminated message")
    return msg, i


def parse(text):
    msg, _ = _parse_msg(_tokens(text), 0, None)
    return msg


def parse_file(path):
    with open(path) as f:
        return parse(f.read())


def get(msg, name, default=None):
    """Last-wins scalar accessor (protobuf optional-field semantics)."""
    v = msg.get(name)
    return v[-1] if v else default
