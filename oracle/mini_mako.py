"""Tiny renderer for the Mako subset used by the reference's ``*-tmpl.cpp`` files.

TEST INFRASTRUCTURE ONLY (oracle build).  The reference build renders three C++
templates with Mako (reference ``setup.py:19-47``); Mako is not installed in this
image and there is no network, so this module interprets the subset those
templates actually use:

* ``<%def name="f(a, b)"> ... </%def>``  -> a Python function returning text
* ``% for ...:`` / ``% if ...:`` / ``% elif`` / ``% else:`` / ``% endfor`` / ``% endif``
  control lines (also written ``%if``)
* ``${expr}`` substitutions, where ``expr`` may call another def

The template is compiled to Python source and executed; nothing here knows
anything about the content of the reference templates.
"""

from __future__ import annotations

import re

_DEF_OPEN = re.compile(r'^\s*<%def\s+name="([^"]+)"\s*>\s*$')
_DEF_CLOSE = re.compile(r"^\s*</%def>\s*$")
_CTRL = re.compile(r"^\s*%\s*(?!%)(.*?)\s*$")
_SUBST = re.compile(r"\$\{(.*?)\}")


def _emit_text(line: str, indent: str) -> str:
    """Python statement appending one text line (with ${} substituted) to __buf."""
    parts = []
    pos = 0
    for m in _SUBST.finditer(line):
        if m.start() > pos:
            parts.append(repr(line[pos : m.start()]))
        parts.append(f"str({m.group(1)})")
        pos = m.end()
    if pos < len(line):
        parts.append(repr(line[pos:]))
    if not parts:
        parts = ["''"]
    return f"{indent}__buf.append({' + '.join(parts)})\n"


def compile_template(src: str) -> str:
    """Translate template text into Python source defining ``__render__()``."""
    out = ["def __render__():\n", "    __buf = []\n"]
    base = "    "
    depth = 0  # control-flow nesting inside the current function
    in_def = False
    for raw in src.splitlines(keepends=True):
        line = raw.rstrip("\n")
        m = _DEF_OPEN.match(line)
        if m:
            if in_def:
                raise ValueError("nested <%def> not supported")
            in_def = True
            out.append(f"    def {m.group(1)}:\n")
            out.append("        __buf = []\n")
            base = "        "
            depth = 0
            continue
        if _DEF_CLOSE.match(line):
            out.append("        return ''.join(__buf)\n")
            in_def = False
            base = "    "
            depth = 0
            continue
        m = _CTRL.match(line)
        if m:
            stmt = m.group(1)
            kw = stmt.split()[0].rstrip(":") if stmt.split() else ""
            if kw in ("endfor", "endif", "endwhile"):
                depth -= 1
            elif kw in ("elif", "else"):
                out.append(f"{base}{'    ' * (depth - 1)}{stmt}\n")
                out.append(f"{base}{'    ' * depth}pass\n")
            elif kw in ("for", "if", "while"):
                out.append(f"{base}{'    ' * depth}{stmt}\n")
                depth += 1
                # guard against empty bodies
                out.append(f"{base}{'    ' * depth}pass\n")
            else:
                raise ValueError(f"unsupported control line: {line!r}")
            continue
        out.append(_emit_text(raw, base + "    " * depth))
    out.append("    return ''.join(__buf)\n")
    return "".join(out)


def render(src: str) -> str:
    ns: dict = {}
    exec(compile(compile_template(src), "<mini_mako>", "exec"), ns)  # noqa: S102
    return ns["__render__"]()


def render_file(path_in: str, path_out: str) -> None:
    with open(path_in) as f:
        txt = render(f.read())
    with open(path_out, "w") as f:
        f.write(txt)
