#!/usr/bin/env python
"""Cut the voxel vertex / fragment shader bodies out of the reference's Samples/SimpleVoxel.cpp into a scratch directory (oracle/_ref/.gen, removed after the compile)
so that ref_driver.cpp can compile and EXECUTE the reference's own shader text (ref_shim/glsl_compat.h supplies the
GLSL vocabulary).  The output is a build product under oracle/_ref/ (git-ignored); nothing from the reference is copied
into the repository.  TEST INFRASTRUCTURE.

Textual changes, all syntax-only:
  * `const vec3 X[56] = vec3[56](` ... `);`   ->  `const vec3 X[56] = {` ... `};`   (GLSL array constructor)
  * `.xyz`                                    ->  `.xyz()`                          (swizzle)
  * `void main()`                             ->  `void vs_main()` / `void fs_main()`
"""
import os
import re
import sys


def cut(src, start_marker, after):
    i = src.index(start_marker, after)
    j = src.index(')";', i)
    return src[i:j], j


def main(ref_root, out_dir):
    src = open(os.path.join(ref_root, "Samples", "SimpleVoxel.cpp"), encoding="utf-8", errors="replace").read()
    vs_at = src.index("class ShaderInstanceVoxelVS")
    fs_at = src.index("class ShaderInstanceVoxelFS")
    vs, _ = cut(src, "ivec3 UnpackU8Vec3(uint PackedValue)", vs_at)
    assert vs_at < src.index("ivec3 UnpackU8Vec3(uint PackedValue)", vs_at) < fs_at
    fs, _ = cut(src, "void main()", fs_at)

    m = re.search(r"=\s*vec3\[56\]\(", vs)
    assert m, "array constructor not found"
    end = vs.index(");", vs.index("//7", m.end()))
    vs = vs[:m.start()] + "= {" + vs[m.end():end] + "};" + vs[end + 2:]
    vs = re.sub(r"\.xyz\b", ".xyz()", vs)
    assert vs.count("void main()") == 1 and fs.count("void main()") == 1
    vs = vs.replace("void main()", "void vs_main()")
    fs = fs.replace("void main()", "void fs_main()")

    os.makedirs(out_dir, exist_ok=True)
    hdr = "/* GENERATED at build time from %s by oracle/extract_ref_glsl.py -- not part of the repository */\n"
    open(os.path.join(out_dir, "ref_voxel_vs.inc"), "w").write(hdr % "Samples/SimpleVoxel.cpp (vertex shader body)" + vs + "\n")
    open(os.path.join(out_dir, "ref_voxel_fs.inc"), "w").write(hdr % "Samples/SimpleVoxel.cpp (fragment shader body)" + fs + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
