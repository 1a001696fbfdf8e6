"""Structure-aware fuzzing of the glTF reader: syntactically valid documents whose fields lie (counts, offsets, strides, component
types, indices into every array, node cycles, deep node chains, sparse / missing members), loaded by the ASan + UBSan build of
tools/fuzz/fuzz_loader.cpp in `files` mode.  python tools/fuzz/craft_gltf.py [n] [seed] [dir]"""
import base64
import json
import random
import struct
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
EXTREME = [-1, 0, 1, 2, 3, 4, 7, 12, 31, 32, 33, 255, 256, 65535, 65536, 2**31 - 1, 2**31, 2**32 - 1, 2**32, 2**53, -2**31, 10**18, 1.5, -0.5, 1e30, None, "x", [], {}, True]


def base_doc(rng):
    import numpy as np
    n = rng.randint(3, 12)
    pos = np.array([[rng.uniform(-1, 1) for _ in range(3)] for _ in range(n)], np.float32)
    nrm = np.array([[0, 0, 1]] * n, np.float32)
    uv = np.array([[rng.random(), rng.random()] for _ in range(n)], np.float32)
    inter = np.concatenate([pos, nrm, uv], 1).astype(np.float32)
    idx_t = rng.choice([(5121, np.uint8), (5123, np.uint16), (5125, np.uint32)])
    ni = 3 * rng.randint(1, 6)
    idx = np.array([rng.randrange(n) for _ in range(ni)], idx_t[1])
    pad = b"\0" * (-(inter.nbytes + idx.nbytes) % 4)
    blob = inter.tobytes() + idx.tobytes() + pad + pos.tobytes()
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"mesh": 0, "children": [1, 2]}, {"mesh": 1, "name": "A"}, {"mesh": 0, "name": "B", "children": []}],
        "materials": [{"name": "m", "pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.5, 0.5, 1], "baseColorTexture": {"index": 0}}, "normalTexture": {"index": 0}}],
        "textures": [{"source": 0}], "images": [{"uri": "albedo.png"}],
        "meshes": [{"name": "a", "primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0, "mode": 4}]},
                   {"name": "b", "primitives": [{"attributes": {"POSITION": 4}}]}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": inter.nbytes, "byteStride": 32},
                        {"buffer": 0, "byteOffset": inter.nbytes, "byteLength": idx.nbytes},
                        {"buffer": 0, "byteOffset": inter.nbytes + idx.nbytes + len(pad), "byteLength": pos.nbytes}],
        "accessors": [{"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": n, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": n, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 24, "componentType": 5126, "count": n, "type": "VEC2"},
                      {"bufferView": 1, "componentType": idx_t[0], "count": ni, "type": "SCALAR"},
                      {"bufferView": 2, "componentType": 5126, "count": n, "type": "VEC3"}],
        "buffers": [{"uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode(), "byteLength": len(blob)}],
    }
    return doc, blob


def paths(node, prefix=()):
    """Every (container, key) in the document."""
    out = []
    if isinstance(node, dict):
        for k, v in node.items():
            out.append((node, k))
            out += paths(v, prefix + (k,))
    elif isinstance(node, list):
        for k, v in enumerate(node):
            out.append((node, k))
            out += paths(v, prefix + (k,))
    return out


def lie(doc, rng):
    for _ in range(rng.randint(1, 3)):
        what = rng.random()
        ps = [(c, k) for c, k in paths(doc) if not (isinstance(c, dict) and k == "uri" and isinstance(c[k], str) and c[k].startswith("data:"))]
        c, k = rng.choice(ps)
        if what < 0.55:
            c[k] = rng.choice(EXTREME)                                  # a field lies
        elif what < 0.7:
            if isinstance(c, dict):
                del c[k]                                                # a member is missing
            else:
                c.pop(k)
        elif what < 0.8:
            nodes = doc.get("nodes")
            dicts = [x for x in nodes if isinstance(x, dict)] if isinstance(nodes, list) else []
            if dicts:
                rng.choice(dicts)["children"] = [rng.randrange(len(nodes)) for _ in range(rng.randint(1, 3))]   # cycles, self references
        elif what < 0.86:
            depth = rng.choice([50, 2000, 50000])                       # a very deep chain of nodes
            doc["nodes"] = [{"children": [i + 1]} for i in range(depth)] + [{"mesh": 0}]
        elif what < 0.92:
            if isinstance(c[k], (int, float)) and not isinstance(c[k], bool):
                c[k] = c[k] + rng.choice([-1, 1, 2, -4, 4, 12, 32])     # off by a little
        else:
            c[k] = [c[k]] * rng.choice([0, 2, 1000])                    # wrong shape


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    out = Path(sys.argv[3] if len(sys.argv) > 3 else "/tmp/cndl_fuzz_gltf")
    out.mkdir(parents=True, exist_ok=True)
    exe = out / "fuzz_loader"
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-x", "c++",
                    str(ROOT / "candela_b200" / "csrc" / "model_loader.cu"), str(ROOT / "tools" / "fuzz" / "fuzz_loader.cpp"), "-o", str(exe)], check=True)
    rng = random.Random(seed)
    files = []
    for i in range(n):
        doc, blob = base_doc(rng)
        if i >= 3:
            lie(doc, rng)
        container = rng.choice(["gltf", "glb"])
        if container == "gltf":
            f = out / f"c{i}.gltf"
            f.write_text(json.dumps(doc))
        else:
            if isinstance(doc.get("buffers"), list) and doc["buffers"] and isinstance(doc["buffers"][0], dict):
                doc["buffers"][0].pop("uri", None)
            js = json.dumps(doc).encode()
            js += b" " * (-len(js) % 4)
            bn = blob + b"\0" * (-len(blob) % 4)
            total = 12 + 8 + len(js) + 8 + len(bn)
            lens = [total, len(js), len(bn)]
            if rng.random() < 0.3:                                       # the container lies too
                lens[rng.randrange(3)] = rng.choice([0, 1, 3, 4, 8, 2**31, 2**32 - 1, total * 2, max(0, len(js) - 4), len(bn) + 4])
            f = out / f"c{i}.glb"
            f.write_bytes(b"glTF" + struct.pack("<II", 2, lens[0] & 0xFFFFFFFF) + struct.pack("<I", lens[1] & 0xFFFFFFFF) + b"JSON" + js +
                          struct.pack("<I", lens[2] & 0xFFFFFFFF) + b"BIN\0" + bn)
        files.append(str(f))
    env = {"ASAN_OPTIONS": "detect_leaks=1:allocator_may_return_null=1:max_allocation_size_mb=2048", "UBSAN_OPTIONS": "print_stacktrace=1"}
    import os
    for lo in range(0, len(files), 500):
        r = subprocess.run([str(exe), "files"] + files[lo:lo + 500], env={**os.environ, **env}, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-6000:])
            sys.exit(r.returncode)
        print(r.stdout.strip())


if __name__ == "__main__":
    main()
