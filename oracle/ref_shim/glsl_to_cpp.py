"""TEST INFRASTRUCTURE ONLY.  Mechanical transliteration of the reference's GLSL traversal include files into C++ that
compiles against the reference's own vendored glm, so that the reference's shader code itself (not a restatement) can be
executed on the CPU and used to pin the oracle:

    python oracle/ref_shim/glsl_to_cpp.py <in.glsl> <out.inc>

Nothing is copied into the repository: the output goes to oracle/_ref/ (git-ignored) and only exists where
/root/reference does.  The rewrites are purely syntactic:
  * `#extension` lines dropped; `layout(...) uniform` / `uniform` declarations become plain globals; every
    `layout (std430, ...) buffer X { T name[]; };` block becomes `static const T* name;`
  * parameter qualifiers: `in` dropped, `out` / `inout` become references
  * unsuffixed floating literals get an `f` (GLSL literals are float; C++ would compute in double)
  * swizzles `.xyz` / `.xy` become `vec3(...)` / `vec2(...)` constructor calls on the same expression
  * the `IntersectRay*` convenience wrappers that pass a swizzle as an `out` argument are dropped (not needed:
    the shim calls IntersectScene* and GetData directly)
  * `vec3(hash2(), hash2().x)` becomes `vec3{hash2(), hash2().x}`: GLSL evaluates constructor arguments left to
    right, C++ leaves the order of function arguments unspecified, and brace initialisation restores the GLSL order

`--function NAME` extracts one top-level function; `NAME@k` takes the k-th definition of an overloaded name (1-based).
"""
import re
import sys


def postfix_start(s: str, dot: int) -> int:
    """Index where the postfix expression ending just before s[dot] == '.' begins."""
    i = dot - 1
    while i >= 0:
        c = s[i]
        if c in ")]":
            close, open_ = c, "(" if c == ")" else "["
            depth = 0
            while i >= 0:
                if s[i] == close:
                    depth += 1
                elif s[i] == open_:
                    depth -= 1
                    if depth == 0:
                        break
                i -= 1
            i -= 1
            # an identifier (function name / array name) may precede the bracket
            while i >= 0 and (s[i].isalnum() or s[i] == "_"):
                i -= 1
            if i >= 0 and s[i] == ".":
                i -= 1
                continue
            break
        elif c.isalnum() or c == "_":
            while i >= 0 and (s[i].isalnum() or s[i] == "_"):
                i -= 1
            if i >= 0 and s[i] == ".":
                i -= 1
                continue
            break
        else:
            break
    return i + 1


def rewrite_swizzles(line: str) -> str:
    for sw, ctor in ((".xyz", "vec3"), (".xy", "vec2")):
        while True:
            m = re.search(re.escape(sw) + r"(?![A-Za-z0-9_])", line)
            if not m:
                break
            st = postfix_start(line, m.start())
            expr = line[st:m.start()]
            c = "u" + ctor if expr.endswith("PackedData") else ctor   # Vertex.PackedData is a uvec4: keep the bits
            line = line[:st] + c + "(" + expr + ")" + line[m.end():]
    return line


def transliterate(text: str) -> str:
    out, lines, i = [], text.splitlines(), 0
    while i < len(lines):
        ln = lines[i]
        st = ln.strip()
        if st.startswith("#extension") or st.startswith("#define SSBO_BINDING_STARTINDEX"):
            i += 1
            continue
        if re.match(r"layout\s*\(\s*std430", st):  # SSBO block -> pointer global
            block = []
            while "};" not in lines[i]:
                block.append(lines[i])
                i += 1
            body = " ".join(block[1:])
            m = re.search(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", body)
            out.append(f"static const {m.group(1)}* {m.group(2)};")
            i += 1
            continue
        if re.match(r"void\s+IntersectRay(IgnoreTransparent)?\s*\(", st):  # wrappers with swizzled out-arguments
            while lines[i].rstrip() != "}":
                i += 1
            i += 1
            continue
        ln = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+", "static ", ln)
        ln = re.sub(r"^\s*uniform\s+", "static ", ln)
        ln = re.sub(r"\b(in\s+const|const\s+in)\s+", "const ", ln)
        ln = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", ln)
        ln = re.sub(r"(?<![\w.])in\s+(?=\w+\s+\w+\s*[,)])", "", ln)
        ln = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", ln)   # 1.0 -> 1.0f ; 1. -> 1.f ; 1.0f untouched
        ln = rewrite_swizzles(ln)
        ln = ln.replace("vec3(hash2(), hash2().x)", "vec3{hash2(), hash2().x}")
        out.append(ln)
        i += 1
    return "\n".join(out) + "\n"


def extract_function(text: str, name: str) -> str:
    """The definition of one top-level function (from its signature line to the closing brace in column 0)."""
    lines = text.splitlines()
    name, _, nth = name.partition("@")
    nth = int(nth or 1)
    for i, ln in enumerate(lines):
        if re.match(r"^\w[\w\s]*\b" + re.escape(name) + r"\s*\(", ln):
            nth -= 1
            if nth:
                continue
            j = i
            while lines[j].rstrip() != "}":
                j += 1
            return "\n".join(lines[i:j + 1]) + "\n"
    raise SystemExit(f"{name} not found")


if __name__ == "__main__":
    if sys.argv[1] == "--function":   # glsl_to_cpp.py --function NAME in.glsl out.inc
        name, src, dst = sys.argv[2], sys.argv[3], sys.argv[4]
        open(dst, "w").write(transliterate(extract_function(open(src, errors="replace").read(), name)))
    else:
        src, dst = sys.argv[1], sys.argv[2]
        open(dst, "w").write(transliterate(open(src, errors="replace").read()))
