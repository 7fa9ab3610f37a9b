"""The three descriptions of the C ABI must agree byte for byte: include/vecchio_gpu.h as gcc lays it out, the ctypes
mirror the Python harness passes through it (vecchio_b200/_abi.py), and the `#[repr(C)]` structs / `extern "C"` block of
rust/gpu.rs -- the binding INTEGRATION.md hands to a maintainer of the reference, which cannot be compiled in this image
(no rustc), so its layouts are computed here by the repr(C) rules and compared with gcc's.

CPU only; no compute call."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vecchio_gpu.h")
RUST = os.path.join(ROOT, "rust", "gpu.rs")


def _strip_c_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _c_structs(text):
    """{name: [field, ...]} in declaration order; an anonymous union counts as one field named '<union>'."""
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{", text):
        name, i, depth = m.group(1), m.end(), 1
        while depth:  # the matching brace (vk_texture nests a union of structs)
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        body = text[m.end():i - 1]
        fields = []
        # an anonymous union: remember its first member as the handle offsetof() can name
        um = re.search(r"union\s*\{(.*)\}\s*;", body, flags=re.S)
        if um:
            first = re.search(r"(\w+)\s*(\[\d+\])*\s*;", um.group(1)).group(1)
            body = body[:um.start()] + f" __union__ {first};" + body[um.end():]
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            if decl.startswith("__union__"):
                fields.append("<union>:" + decl.split()[1])
                continue
            # "type a[3], b, c": the type is everything up to the last space of the first declarator; pointers carry their
            # star on the type ("const vk_node* nodes"), "const float* a, *b" does not occur in the header
            declarators = decl.split(",")
            names = [declarators[0].rpartition(" ")[2]] + [x.strip() for x in declarators[1:]]
            for n in names:
                fields.append(re.sub(r"\[.*", "", n).lstrip("*"))
        out[name] = fields
    return out


def _c_prototypes(text):
    """{function: number of parameters}"""
    protos = {}
    for m in re.finditer(r"\b(vk_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


@pytest.fixture(scope="module")
def c_layout(tmp_path_factory):
    """sizeof / offsetof of every struct of the header, by compiling a C11 program against it."""
    text = _strip_c_comments(open(HEADER).read())
    structs = _c_structs(text)
    assert {"vk_node", "vk_scene_desc", "vk_camera", "vk_render_params", "vk_stats", "vk_eval", "vk_texture"} <= set(structs)
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for s, fields in structs.items():
        lines.append(f'  printf("S {s} %zu\\n", sizeof({s}));')
        for f in fields:
            label, member = (f.split(":")[0], f.split(":")[1]) if f.startswith("<union>") else (f, f)
            lines.append(f'  printf("F {s} {label} %zu %zu\\n", offsetof({s}, {member}), sizeof((({s}*)0)->{member}));')
    lines += ["  return 0;", "}"]
    d = tmp_path_factory.mktemp("abi")
    src, exe = d / "layout.c", d / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    layout = {}
    for line in out.splitlines():
        p = line.split()
        if p[0] == "S":
            layout[p[1]] = {"size": int(p[2]), "fields": []}
        else:
            layout[p[1]]["fields"].append((p[2], int(p[3]), int(p[4])))
    return layout, _c_prototypes(text)


# ---------------------------------------------------------------------------------------------------------------------
# repr(C) layout of rust/gpu.rs, computed (the same natural-alignment rules gcc applies on x86-64 / aarch64 Linux)

_RUST_PRIM = {"f32": (4, 4), "u32": (4, 4), "i32": (4, 4), "u64": (8, 8), "i64": (8, 8), "u8": (1, 1), "i8": (1, 1), "u16": (2, 2),
              "usize": (8, 8), "isize": (8, 8), "c_int": (4, 4), "c_uint": (4, 4), "c_char": (1, 1), "vk_ref": (4, 4), "f64": (8, 8)}


def _rust_type_layout(t, structs):
    t = t.strip()
    if t.startswith("*const") or t.startswith("*mut"):
        return 8, 8
    m = re.fullmatch(r"\[(.+);\s*(\d+)\]", t)
    if m:
        size, align = _rust_type_layout(m.group(1), structs)
        return size * int(m.group(2)), align
    if t in _RUST_PRIM:
        return _RUST_PRIM[t]
    if t in structs:
        return structs[t]["size"], structs[t]["align"]
    raise AssertionError(f"rust/gpu.rs: type `{t}` is not one this test knows how to lay out")


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        depth += ch in "[(<"
        depth -= ch in "])>"
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def _rust_layout():
    text = re.sub(r"//[^\n]*", " ", open(RUST).read())
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\](?:\s*#\[[^\]]*\])*\s*pub\s+struct\s+(\w+)\s*\{(.*?)\}", text, flags=re.S):
        name, body = m.group(1), m.group(2)
        off, align, fields = 0, 1, []
        for part in _split_top_level(body):
            fname, _, ftype = part.strip().removeprefix("pub ").partition(":")
            size, a = _rust_type_layout(ftype, structs)
            off = (off + a - 1) // a * a
            fields.append((fname.strip(), off, size))
            off += size
            align = max(align, a)
        structs[name] = {"size": (off + align - 1) // align * align, "align": align, "fields": fields}
    protos = {}
    for block in re.finditer(r'extern\s+"C"\s*\{(.*?)\n\}', text, flags=re.S):
        for m in re.finditer(r"pub\s+fn\s+(vk_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block.group(1), flags=re.S):
            args = m.group(2).strip()
            protos[m.group(1)] = len(_split_top_level(args)) if args else 0
    return structs, protos


def test_ctypes_mirror_has_the_layout_gcc_gives_the_header(c_layout):
    import vecchio_b200._abi as abi
    layout, _ = c_layout
    checked = 0
    for name, want in layout.items():
        cls = getattr(abi, name, None)
        if cls is None:
            continue
        assert C.sizeof(cls) == want["size"], f"{name}: ctypes {C.sizeof(cls)} B, C {want['size']} B"
        mirror = {n: getattr(cls, n) for n, *_ in cls._fields_}
        for fname, off, size in want["fields"]:
            if fname == "<union>":
                continue  # vk_texture: the mirror views the union as three words; offset checked through the size above
            if fname not in mirror:
                continue  # a mirror may fold consecutive fields (checked by size); named ones must sit where C puts them
            assert (mirror[fname].offset, mirror[fname].size) == (off, size), f"{name}.{fname}"
            checked += 1
    # every struct a call passes by pointer has a mirror
    for name in ("vk_node", "vk_sphere", "vk_msphere", "vk_rect", "vk_box", "vk_xform", "vk_medium", "vk_material", "vk_texture",
                 "vk_perlin", "vk_scene_desc", "vk_camera", "vk_render_params", "vk_stats", "vk_ray", "vk_hit", "vk_scene_info"):
        assert hasattr(abi, name), f"_abi.py has no mirror of {name}"
        assert C.sizeof(getattr(abi, name)) == layout[name]["size"], name
    assert checked > 100
    # the batch records the tests build with numpy
    for name, dt in (("vk_ray", abi.RAY_DTYPE), ("vk_hit", abi.HIT_DTYPE), ("vk_eval", abi.EVAL_DTYPE)):
        assert dt.itemsize == layout[name]["size"], name
        assert [n for n, *_ in layout[name]["fields"]] == list(dt.names), f"{name}: field order"
        for fname, off, size in layout[name]["fields"]:
            sub, at = dt.fields[fname][:2]
            assert (at, sub.itemsize) == (off, size), f"{name}.{fname}"


def test_rust_binding_structs_have_the_layout_gcc_gives_the_header(c_layout):
    layout, _ = c_layout
    rust, _ = _rust_layout()
    # the structs the binding passes through the calls it declares
    for name in ("vk_node", "vk_sphere", "vk_msphere", "vk_rect", "vk_box", "vk_xform", "vk_medium", "vk_material", "vk_texture",
                 "vk_perlin", "vk_scene_desc", "vk_camera", "vk_render_params", "vk_stats"):
        assert name in rust, f"rust/gpu.rs has no #[repr(C)] struct {name}"
        want, got = layout[name], rust[name]
        assert got["size"] == want["size"], f"{name}: rust {got['size']} B, C {want['size']} B"
        assert len(got["fields"]) == len(want["fields"]), f"{name}: field count"
        for (rn, roff, rsize), (cn, coff, csize) in zip(got["fields"], want["fields"]):
            if cn == "<union>":
                assert roff == coff and rsize == want["size"] - coff, f"{name}: union payload"
                continue
            assert rn.rstrip("_") == cn, f"{name}: field order, rust `{rn}` against C `{cn}`"
            assert (roff, rsize) == (coff, csize), f"{name}.{cn}: rust at {roff}+{rsize}, C at {coff}+{csize}"
    # the record sizes the header's comments promise, the kernels' 128-bit loads rely on them
    for name, size in (("vk_node", 32), ("vk_sphere", 16), ("vk_msphere", 48), ("vk_rect", 32), ("vk_box", 32), ("vk_xform", 32),
                       ("vk_medium", 16), ("vk_material", 16), ("vk_texture", 16), ("vk_camera", 96), ("vk_ray", 36), ("vk_hit", 56),
                       ("vk_eval", 43 * 4)):
        assert layout[name]["size"] == size, name


def test_rust_binding_declares_the_header_s_functions_with_the_same_arity(c_layout):
    _, c_protos = c_layout
    _, rust_protos = _rust_layout()
    assert rust_protos, "no extern \"C\" block found in rust/gpu.rs"
    for fn, n in rust_protos.items():
        assert fn in c_protos, f"rust/gpu.rs declares {fn}, include/vecchio_gpu.h does not"
        assert n == c_protos[fn], f"{fn}: {n} parameters in rust/gpu.rs, {c_protos[fn]} in the header"
    # what the binding needs to replace the sample loop on one or several GPUs
    for fn in ("vk_create", "vk_destroy", "vk_last_error", "vk_scene_upload", "vk_render", "vk_render_rgb8", "vk_multi_create",
               "vk_multi_destroy", "vk_multi_scene_upload", "vk_multi_render", "vk_multi_render_rgb8"):
        assert fn in rust_protos, f"rust/gpu.rs does not declare {fn}"


def test_rust_binding_constants_equal_the_header_s():
    text = _strip_c_comments(open(HEADER).read())
    rust = open(RUST).read()
    consts = dict(re.findall(r"pub const (VK_\w+): \w+ = ([^;]+);", rust))
    assert consts
    enums = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", text, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            k, _, v = item.partition("=")
            try:
                nxt = int(v.strip().rstrip("u"), 0) if v.strip() else nxt
            except ValueError:
                continue  # an expression; none of the constants the binding copies is one
            enums[k.strip()] = nxt
            nxt += 1
    enums.update({k: int(v.rstrip("uU"), 0) for k, v in re.findall(r"#define\s+(VK_\w+)\s+(0x[0-9a-fA-F]+u?|\d+u?)\s*$", text, flags=re.M)})
    for k, v in consts.items():
        assert k in enums, f"{k} is not a constant of the header"
        assert int(v.replace("_", ""), 0) == enums[k], f"{k}: rust {v}, header {enums[k]}"


def test_rust_lower_impls_use_what_the_binding_defines():
    """rust/lower_impls.rs (the lower() of every shipped type, uncompiled like gpu.rs): every `gpu::` name it uses is defined
    in rust/gpu.rs, every struct literal names exactly the struct's fields, every reference type of SURVEY 8(a) has a block."""
    rust = re.sub(r"//[^\n]*", " ", open(RUST).read())
    impls_raw = open(os.path.join(ROOT, "rust", "lower_impls.rs")).read()
    impls = re.sub(r"//[^\n]*", " ", impls_raw)
    defined = set(re.findall(r"pub\s+(?:const\s+fn|const|fn|struct|enum|type)\s+(\w+)", rust))
    used = set(re.findall(r"\bgpu::(\w+)", impls))
    assert used and used <= defined, f"not defined in rust/gpu.rs: {sorted(used - defined)}"
    methods = set(re.findall(r"pub\s+fn\s+(\w+)\s*\(\s*&mut\s+self", rust))
    for m in set(re.findall(r"\bb\.(\w+)\(", impls)):
        assert m in methods, f"Lowering has no method `{m}`"
    structs, _ = _rust_layout()
    n_literals = 0
    for name, body in re.findall(r"gpu::(vk_\w+)\s*\{([^{}]*)\}", impls):
        fields = [part.split(":")[0].strip() for part in _split_top_level(body) if part.strip()]
        assert sorted(fields) == sorted(f for f, *_ in structs[name]["fields"]), f"{name} literal: fields {fields}"
        n_literals += 1
    assert n_literals >= 8
    vec_fields = set(re.findall(r"pub\s+(\w+):\s*Vec<", rust)) | set(re.findall(r"pub\s+(memo_\w+):", rust))
    for f in set(re.findall(r"\bb\.(\w+)\.(?:push|len|get|insert|extend_from_slice)\b", impls)):
        assert f in vec_fields, f"Lowering has no field `{f}`"
    for ty in ("Sphere", "MovingSphere", "Rect", "FlipFace", "Boxy", "ConstantMedium", "Translate", "RotateX", "RotateY", "RotateZ", "BVHNode",
               "Lambertian", "Metal", "Dielectric", "DiffuseLight", "Isotropic", "SpecDiffuse", "SolidColor", "Checker", "ImageTexture",
               "NoiseTexture", "Camera"):
        assert re.search(rf"^impl {ty} \{{", impls_raw, flags=re.M), f"no lower() block for {ty}"
