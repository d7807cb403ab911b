"""Host mirror of numcl's einsum front end (src/3einsum.lisp): the subscript reader and normaliser.

In a Lisp deployment this code is not needed -- numcl's own `einsum-normalize-subscripts`
produces the `einsum-specs` struct and the `:cuda` method of `einsum-body` serialises it
(lisp/cuda.lisp).  The Python host mirror restates it so that the same subscripts can be written in
tests:  einsum("ij jk -> ik", a, b)  <->  (einsum '(ij jk -> ik) a b).

Restated functions (file:line in the numcl tree):
  einsum-normalize-subscripts   src/3einsum.lisp:64-96
  %einsum-normalize-subscripts  src/3einsum.lisp:98-200
  sort-locality                 src/3einsum.lisp:523-565
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field


class Sym(str):
    """A Lisp symbol (upcased by the reader).  Keywords keep a leading colon."""
    __slots__ = ()

    def __repr__(self):
        return self.lower()


class EinsumSpecError(TypeError):
    """type-error / assertion failures of the reference's normaliser."""


_FLOAT_RE = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eEfFdD][+-]?\d+)?$")


def _atom(tok: str):
    if re.fullmatch(r"[+-]?\d+", tok):
        return int(tok)
    if _FLOAT_RE.match(tok) and any(c in tok for c in ".eEfFdD") and any(c.isdigit() for c in tok):
        if tok[-1] == ".":      # "10." reads as the integer 10
            return int(tok[:-1])
        m = re.search(r"[dD]", tok)
        val = float(re.sub(r"[eEfFdD]", "e", tok))
        return Double(val) if m else val
    if tok.startswith(":"):
        return Sym(":" + tok[1:].upper())
    if ":" in tok:              # package prefix: cl:+ -> +
        tok = tok.rsplit(":", 1)[1]
    return Sym(tok.upper())


class Double(float):
    """A double-float literal (1.0d0); plain Python floats read as single-floats."""
    __slots__ = ()


def read_subscripts(text: str) -> list:
    """Read the contents of the subscripts list, e.g. "ij jk -> ik" for '(ij jk -> ik)."""
    toks = re.findall(r"\(|\)|'|[^\s()']+", text)
    pos = 0

    def rd_list(top):
        nonlocal pos
        out = []
        while pos < len(toks):
            t = toks[pos]
            pos += 1
            if t == "(":
                out.append(rd_list(False))
            elif t == ")":
                if top:
                    raise EinsumSpecError("unbalanced )")
                return out
            elif t == "'":
                continue
            else:
                out.append(_atom(t))
        if not top:
            raise EinsumSpecError("unbalanced (")
        return out

    return rd_list(True)


@dataclass
class IterOpt:
    key: object                     # index number after sublis (or the untouched symbol)
    start: int | None = None        # None = default (0)
    end: int | None = None          # None = default (array-dimension)
    step: int = 1


@dataclass
class EinsumSpecs:
    """einsum-specs, src/3einsum.lisp:44-62."""
    iter_specs: list
    transforms: list
    i_specs: list
    o_specs: list
    i_options: list = field(default_factory=list)
    o_options: list = field(default_factory=list)

    def describe(self) -> str:
        def ilist(v):
            return "(" + " ".join(str(x) for x in v) + ")"

        def form(f):
            if isinstance(f, list):
                return "(" + " ".join(form(x) for x in f) + ")"
            if isinstance(f, Sym):
                return f.lower()
            if isinstance(f, Double):
                return "%.17gd0" % float(f)
            if isinstance(f, float):
                return "%.9g" % f
            return repr(f)

        def opts(o):
            if not o:
                return "nil"
            parts = []
            for e in o:
                s = f"({e.key if not isinstance(e.key, Sym) else e.key.lower()}"
                if e.start is not None:
                    s += f" :start {e.start}"
                if e.end is not None:
                    s += f" :end {e.end}"
                if e.step != 1:
                    s += f" :step {e.step}"
                parts.append(s + ")")
            return "(" + " ".join(parts) + ")"

        return (f"iter={ilist(self.iter_specs)} i=({' '.join(ilist(s) for s in self.i_specs)})"
                f" o=({' '.join(ilist(s) for s in self.o_specs)}) t=({' '.join(form(t) for t in self.transforms)})"
                f" io=({' '.join(opts(o) for o in self.i_options)}) oo=({' '.join(opts(o) for o in self.o_options)})")


def _is_arrow(x) -> bool:
    return isinstance(x, Sym) and x == "->"


def _explode(s) -> list:
    """explode, src/3einsum.lisp:124-138."""
    if isinstance(s, list):
        for c in s:
            if not isinstance(c, Sym):
                raise EinsumSpecError("(assert (symbolp c)) failed")
        return [Sym(c) for c in s]
    if isinstance(s, Sym):
        out = []
        for c in str(s):
            if not (c == "-" or c.isalpha()):
                raise EinsumSpecError(f"Tried to make a spec from a non-alpha char {c}")
            out.append(Sym(c))
        return out
    raise EinsumSpecError(f"spec {s!r} is neither a symbol nor a list")


def sort_locality(indices: list, subscripts: list) -> list:
    """sort-locality, src/3einsum.lisp:523-565: greedy earliest-position order with a penalty for
    taking the next index from the same set of subscripts as an earlier one."""
    indices = list(indices)
    sticky: dict = {}

    def key_of(index):
        return tuple(i for i, sub in enumerate(subscripts) if index in sub)

    def score(index):
        return sum(sub.index(index) + 1 for sub in subscripts if index in sub)

    order = []
    while indices:
        best, best_score = None, None
        for index in indices:                      # `finding ... minimizing`: first minimum wins
            s = score(index) + len(indices) * sticky.get(key_of(index), 0)
            if best_score is None or s < best_score:
                best, best_score = index, s
        k = key_of(best)
        sticky[k] = sticky.get(k, 0) + 1
        order.append(best)
        indices.remove(best)
    return order


def _remove_duplicates(seq):
    """cl:remove-duplicates keeps the LAST occurrence of each element."""
    out = []
    for i, x in enumerate(seq):
        if x not in seq[i + 1:]:
            out.append(x)
    return out


def _sbcl_union(l1, l2):
    """cl:union as SBCL computes it for short lists: members of l1 missing from l2 are pushed onto
    l2 [SBCL, unverified here; only the tie order of sort-locality depends on it]."""
    res = list(l2)
    for e in l1:
        if e not in l2:
            res.insert(0, e)
    return res


def _parse_options(alist, mapping) -> list:
    out = []
    if not alist or alist == Sym("NIL"):
        return out
    if not isinstance(alist, list):
        raise EinsumSpecError("iteration options must be an alist")
    for ent in alist:
        if not isinstance(ent, list) or not ent:
            raise EinsumSpecError("bad iteration option entry")
        key = ent[0]
        if isinstance(key, Sym) and key in mapping:          # (sublis alist options): symbol -> number
            key = mapping[key]
        o = IterOpt(key=key)
        for k in range(1, len(ent) - 1, 2):
            kw, v = ent[k], ent[k + 1]
            is_t = isinstance(v, Sym) and v == "T"
            if not is_t and not isinstance(v, int):
                raise EinsumSpecError(f"{kw} = {v}: iteration specifier option should be constant")
            if kw == ":START" and not is_t:
                o.start = v
            elif kw == ":END" and not is_t:
                o.end = v
            elif kw == ":STEP":
                o.step = v
        out.append(o)
    return out


def einsum_normalize_subscripts(subscripts) -> EinsumSpecs:
    """einsum-normalize-subscripts + %einsum-normalize-subscripts, src/3einsum.lisp:64-200."""
    if isinstance(subscripts, str):
        subscripts = read_subscripts(subscripts)
    arrows = [i for i, x in enumerate(subscripts) if _is_arrow(x)]
    if len(arrows) > 4:
        raise EinsumSpecError("ecase: too many ->")
    segs, b = [], 0
    for a in arrows + [len(subscripts)]:
        segs.append(subscripts[b:a])
        b = a + 1
    na = len(arrows)
    i_raw = segs[0]
    t_raw = segs[1] if na >= 2 else None
    o_raw = segs[2] if na >= 2 else (segs[1] if na == 1 else None)
    io_raw = segs[3] if na >= 3 else []
    oo_raw = segs[4] if na >= 4 else []
    if na >= 1 and not o_raw:
        o_raw = [[]]                                   # (or (subseq ...) '(()))

    i_specs = [_explode(s) for s in i_raw]
    indices = []
    for spec in i_specs:                               # pushnew ... nreverse: order of first appearance
        for ix in spec:
            if ix not in indices:
                indices.append(ix)
    if o_raw is None:
        o_specs = [sorted(indices, key=str)]           # (sort (copy-list indices) #'string<)
    else:
        o_specs = [_explode(s) for s in o_raw]
    mapping = {ix: (-1 if ix == "-" else i) for i, ix in enumerate(indices)}   # make-map
    i_num = [[mapping[ix] for ix in spec] for spec in i_specs]
    o_num = []
    for spec in o_specs:
        row = []
        for ix in spec:
            if ix not in mapping:
                raise EinsumSpecError(f"The output spec contains {ix!r} which are not used in the input specs")
            row.append(mapping[ix])
        o_num.append(row)

    i_len, o_len = len(i_specs), len(o_specs)
    if t_raw:
        transforms = list(t_raw)
    else:                                              # `(+ ,(@ o) (* ,@(mapcar #'$ (iota i-len :start 1))))
        transforms = [[Sym("+"), Sym(f"@{o}"), [Sym("*")] + [Sym(f"${i}") for i in range(1, i_len + 1)]]
                      for o in range(1, o_len + 1)]
    if len(transforms) != o_len:
        raise EinsumSpecError("(assert (= o-len (length transforms))) failed")

    i_options = [[] for _ in range(i_len)]
    o_options = [[] for _ in range(o_len)]
    for k, alist in enumerate(io_raw[:i_len]):         # (replace full given) stops at the shorter
        i_options[k] = _parse_options(alist, mapping)
    for k, alist in enumerate(oo_raw[:o_len]):
        o_options[k] = _parse_options(alist, mapping)

    i_flat = _remove_duplicates([x for s in i_num for x in s])
    o_flat = _remove_duplicates([x for s in o_num for x in s])
    iter_specs = sort_locality(_sbcl_union(i_flat, o_flat), i_num + o_num)
    return EinsumSpecs(iter_specs=iter_specs, transforms=transforms, i_specs=i_num, o_specs=o_num,
                       i_options=i_options, o_options=o_options)
