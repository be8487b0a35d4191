"""The Java FFM bindings under java/ cannot be compiled in this image (no JDK).  What can be checked without one: every downcall handle
`h("jl_xxx", FunctionDescriptor.of(...))` must describe the prototype include/jlama_b200.h declares for that symbol -- same arity, same
machine type per argument and return value (ADDRESS for pointers, JAVA_INT / JAVA_LONG / JAVA_FLOAT / JAVA_DOUBLE for scalars) -- and the
jl_model_config StructLayout must list the C struct's members in order.  A wrong descriptor is the classic FFM failure: it links, and
then passes garbage."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIND = {"int": "JAVA_INT", "int32_t": "JAVA_INT", "int64_t": "JAVA_LONG", "uint64_t": "JAVA_LONG", "float": "JAVA_FLOAT", "double": "JAVA_DOUBLE"}


def _header():
    hdr = open(os.path.join(ROOT, "include", "jlama_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return hdr


def _c_layout(decl):
    d = decl.strip()
    if "*" in d or "[" in d:
        return "ADDRESS"
    words = [w for w in re.split(r"\s+", d) if w not in ("const", "unsigned", "struct")]
    return KIND[words[0]]


def _prototypes():
    hdr = re.sub(r"typedef struct \{[^{}]*\} \w+;", "", _header(), flags=re.S)
    hdr = re.sub(r"^\s*#.*$", "", hdr, flags=re.M)
    out = {}
    for ret, name, params in re.findall(r"([A-Za-z_][\w\s\*]*?)\b(jl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        plist = [] if params.strip() in ("", "void") else params.split(",")
        out[name] = [_c_layout(ret + " x")] + [_c_layout(p) for p in plist]
    return out


def test_every_java_downcall_descriptor_matches_its_c_prototype():
    protos = _prototypes()
    checked = 0
    for path in glob.glob(os.path.join(ROOT, "java", "**", "*.java"), recursive=True):
        src = re.sub(r"//.*$", "", open(path).read(), flags=re.M)
        for name, layouts in re.findall(r'h\(\s*"(jl_\w+)"\s*,\s*FunctionDescriptor\.of\(([^;]*?)\)\s*\)\s*;', src, flags=re.S):
            got = [x.strip() for x in layouts.split(",")]
            assert name in protos, (os.path.basename(path), name, "not declared in include/jlama_b200.h")
            assert got == protos[name], (os.path.basename(path), name, got, protos[name])
            checked += 1
    n_handles = sum(len(re.findall(r'h\(\s*"jl_\w+"', open(p).read())) for p in glob.glob(os.path.join(ROOT, "java", "**", "*.java"), recursive=True))
    assert checked == n_handles >= 30  # every handle in the sources was parsed and compared


def test_java_model_config_layout_lists_the_c_members_in_order():
    body = dict((n, b) for b, n in re.findall(r"typedef struct \{([^{}]*)\} (\w+);", _header()))["jl_model_config"]
    members = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ctype = stmt.split()[0]
        for nm in stmt[len(ctype):].split(","):
            members.append((nm.strip(), KIND[ctype]))
    src = open(os.path.join(ROOT, "java", "com", "github", "tjake", "jlama", "model", "CudaLlamaModel.java")).read()
    layout = re.search(r"StructLayout CONFIG = MemoryLayout\.structLayout\((.*?)\);", src, flags=re.S).group(1)
    layout = re.sub(r"/\*.*?\*/", "", layout, flags=re.S)
    java = re.findall(r'(JAVA_\w+)\.withName\("(\w+)"\)|MemoryLayout\.(paddingLayout)\((\d+)\)', layout)
    got = [(n, k) for k, n, pad, _ in java if not pad]
    assert got == members, (got, members)
    # natural alignment: the one padding entry sits before the first double, after an odd number of 4-byte members
    seq = [("pad" if pad else k) for k, n, pad, _ in java]
    first_double = seq.index("JAVA_DOUBLE")
    assert seq[first_double - 1] == "pad" and (first_double - 1) % 2 == 1
    assert seq.count("pad") == 1
    import ctypes as C
    from jlama_b200 import native
    size = sum(8 if k == "JAVA_DOUBLE" else 4 for k in seq)
    assert size == C.sizeof(native.ModelConfig)


def _top_level_args(src, open_idx):
    """arguments of the call whose '(' is at open_idx, split at depth-0 commas"""
    depth, i, start, args = 0, open_idx, open_idx + 1, []
    while True:
        c = src[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                tail = src[start:i].strip()
                if tail:
                    args.append(tail)
                return args
        elif c == "," and depth == 1:
            args.append(src[start:i].strip())
            start = i + 1
        i += 1


def test_every_invoke_exact_passes_what_its_descriptor_declares():
    """invokeExact is checked against the handle's type only at run time (WrongMethodTypeException): count the arguments of every call
    site and compare the cast of the result with the descriptor's return layout."""
    protos = _prototypes()
    ret_cast = {"JAVA_INT": "int", "JAVA_LONG": "long", "ADDRESS": "MemorySegment", "JAVA_FLOAT": "float", "JAVA_DOUBLE": "double"}
    calls = 0
    for path in glob.glob(os.path.join(ROOT, "java", "**", "*.java"), recursive=True):
        src = re.sub(r"//.*$", "", open(path).read(), flags=re.M)
        for m in re.finditer(r"\(\s*(\w+)\s*\)\s*(jl_\w+)\.invokeExact\(", src):
            cast, name = m.group(1), m.group(2)
            args = _top_level_args(src, m.end() - 1)
            want = protos[name]
            assert len(args) == len(want) - 1, (os.path.basename(path), name, len(args), len(want) - 1)
            assert cast == ret_cast[want[0]], (os.path.basename(path), name, cast, want[0])
            # literal arguments must have the Java type the layout implies (an int literal where a long is declared fails at run time)
            for a, layout in zip(args, want[1:]):
                if re.fullmatch(r"-?\d+", a):
                    assert layout == "JAVA_INT", (os.path.basename(path), name, a, layout)
                if re.fullmatch(r"-?\d+L", a):
                    assert layout == "JAVA_LONG", (os.path.basename(path), name, a, layout)
            calls += 1
    assert calls >= 30
