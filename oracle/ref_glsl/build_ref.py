#!/usr/bin/env python
"""Builds oracle/_ref/libasvgf_ref.so: the reference's OWN compute shaders, compiled as C++.  TEST INFRASTRUCTURE ONLY.

The reference implements the denoiser path as GLSL compute shaders (/root/reference/Source/Shaders/*.comp) that its
build compiles with glslc and runs through Vulkan; neither exists in this image.  The shader text itself is plain
GLSL arithmetic, so this script compiles it for the CPU instead:

  1. the UNMODIFIED shader file is read where it lies under /root/reference, its #include chain is flattened and the
     `#version` / `#extension` lines dropped; g++ -E then resolves every #define / #if exactly as glslc would;
  2. what is GLSL syntax but not C++ is rewritten mechanically (no expression of the shader is touched):
       layout(local_size_*) in;                      -> constants for the dispatcher
       layout(constant_id = N) const uint x = d;     -> const uint x = specConstant(N)
       layout(set, binding[, fmt]) uniform image2D n -> image2D n = bindImage(binding, fmt)   (same for samplers)
       layout(set, binding) uniform Block { T m; };  -> const T& m = bound buffer
       shared T v[..];                               -> thread_local static (one OS thread runs one work group)
       out / inout parameters                        -> C++ references
       1.0                                           -> 1.0f   (GLSL literals are single precision)
       file-scope `const T x = ...;`                 -> static constexpr members (unless they use a specialisation constant)
  3. the result becomes the body of `struct Shader : glslshim::InvocationBase` -- one object per invocation -- and is
     compiled with the reference's vendored GLM (Tests/Libs/glm) for the vector types, over oracle/ref_glsl/glsl_shim.h.

Nothing is copied into the repository: the generated C++ and the .so go to oracle/_ref/ (git-ignored).
Usage: python oracle/ref_glsl/build_ref.py [--reference /root/reference] [--keep]
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")

# shader file -> does it call barrier()
SHADERS = [
    "CmASVGFGradientAtrous.comp",
    "CmSVGFTemporalAccumulation.comp",
    "CmAntiFirefly.comp",
    "CmSVGFEstimateVariance.comp",
    "CmSVGFAtrous_Iter0.comp",
    "CmSVGFAtrous.comp",
    "CmCheckerboard.comp",
    "CmLuminanceHistogram.comp",
    "CmLuminanceAvg.comp",
    "CmScatterAccum.comp",
    "CmPrepareFinal.comp",
]

FORMATS = ["r8ui", "r32ui", "rgba32ui", "rg8", "rgba8", "r8", "r16f", "rg16f", "rgba16f", "r32f", "rgba32f", "r11f_g11f_b10f"]


def flatten(path, dirs, seen):
    """The text of `path` with its #include "..." lines replaced by the included files (each file once)."""
    out = []
    for line in open(path, encoding="utf-8", errors="replace").read().splitlines():
        m = re.match(r'\s*#\s*include\s+"([^"]+)"', line)
        if m:
            for d in [os.path.dirname(path)] + dirs:
                cand = os.path.normpath(os.path.join(d, m.group(1)))
                if os.path.exists(cand):
                    if cand not in seen:
                        seen.add(cand)
                        out.append(flatten(cand, dirs, seen))
                    break
            else:
                raise FileNotFoundError("%s includes %s" % (path, m.group(1)))
            continue
        if re.match(r"\s*#\s*(version|extension)\b", line):
            out.append("")
            continue
        out.append(line)
    return "\n".join(out)


def preprocess(text):
    r = subprocess.run(["g++", "-E", "-P", "-x", "c++", "-undef", "-nostdinc", "-w", "-"], input=text, capture_output=True,
                       text=True)
    if r.returncode != 0:
        raise RuntimeError("cpp failed:\n" + r.stderr[-4000:])
    return r.stdout


FLOAT_LITERAL = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def split_top_level(text):
    """Yields (chunk, depth_before) where chunks end at ';' or '}' of brace depth 0 -- enough to find file-scope statements."""
    depth, start, chunks = 0, 0, []
    i = 0
    while i < len(text):
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                # a function body or a struct/initialiser: include a directly following ';'
                j = i + 1
                while j < len(text) and text[j] in " \t\r\n":
                    j += 1
                if j < len(text) and text[j] == ";":
                    i = j
                chunks.append(text[start:i + 1])
                start = i + 1
        elif c == ";" and depth == 0:
            chunks.append(text[start:i + 1])
            start = i + 1
        i += 1
    if text[start:].strip():
        chunks.append(text[start:])
    return chunks


def transform(glsl, name):
    info = {"local": None, "barrier": bool(re.search(r"\b(barrier|groupMemoryBarrier)\s*\(", glsl)), "spec": [], "buffers": [], "push": False}
    t = glsl

    spec_names = []
    out, pos = [], 0
    for m in re.finditer(r"\blayout\s*\(", t):
        if m.start() < pos:
            continue
        # balanced qualifier list
        depth, j = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(t[j], 0)
            j += 1
        quals = t[m.end():j - 1]
        q = {}
        for item in re.split(r",(?![^()]*\))", quals):
            k, _, v = item.partition("=")
            q[k.strip()] = v.strip()
        rest = t[j:]
        out.append(t[pos:m.start()])
        mm = re.match(r"\s*in\s*;", rest)
        if mm and "local_size_x" in q:
            info["local"] = (q["local_size_x"], q.get("local_size_y", "1"), q.get("local_size_z", "1"))
            pos = j + mm.end()
            continue
        mm = re.match(r"\s*const\s+(\w+)\s+(\w+)\s*=\s*([^;]+);", rest)
        if mm and "constant_id" in q:
            spec_names.append(mm.group(2))
            info["spec"].append((q["constant_id"], mm.group(2)))
            out.append("const %s %s = (%s)glslshim::specConstant(%s, %s);" % (mm.group(1), mm.group(2), mm.group(1), q["constant_id"], mm.group(3).strip()))
            pos = j + mm.end()
            continue
        mm = re.match(r"\s*(?:readonly\s+|writeonly\s+|restrict\s+|coherent\s+)*uniform\s+(u?image[23]D|u?sampler[23]D|texture2DArray|sampler2DArray|samplerCube)\s+(\w+)\s*;", rest)
        if mm and not mm.group(1).endswith("2D"):   # the volumetric images: declared by some headers, never sampled on this path
            out.append("const %s %s = {};" % (mm.group(1), mm.group(2)))
            pos = j + mm.end()
            continue
        if mm:
            fmt = [k for k in q if k in FORMATS]
            other = [k for k in q if k not in FORMATS and k not in ("set", "binding")]
            if other:
                raise RuntimeError("%s: image %s has qualifiers %s" % (name, mm.group(2), other))
            out.append("const %s %s = glslshim::bindImage< glslshim::%s >(%s, %s, glslshim::%s);" %
                       (mm.group(1), mm.group(2), mm.group(1), q["set"], q["binding"], "FMT_" + fmt[0] if fmt else "FMT_NONE"))
            pos = j + mm.end()
            continue
        mm = re.match(r"\s*(?:readonly\s+|writeonly\s+|restrict\s+|coherent\s+)*(uniform|buffer)\s+(\w+)\s*\{([^}]*)\}\s*(\w*)\s*;", rest)
        if mm:
            kind, bname, body, inst = mm.groups()
            members = [x.strip() for x in body.split(";") if x.strip()]
            if "push_constant" in q:
                info["push"] = True
                src = "glslshim::g_ctx->pushConstant"
            elif kind == "uniform":
                src = "glslshim::uniformBlock(%s, %s)" % (q["set"], q["binding"])
            else:
                src = "glslshim::storageBlock(%s, %s)" % (q["set"], q["binding"])
                info["buffers"].append(bname)
            if inst:  # instance name: members are reached as inst.member -> a struct of that layout
                out.append("struct %s_T { %s; };\n%s_T& %s = *( %s_T* )( %s );" % (bname, "; ".join(members), bname, inst, bname, src))
            else:
                for mem in members:
                    m2 = re.match(r"(\w+)\s+(\w+)(?:\s*\[\s*(\w*)\s*\])?$", mem)
                    if not m2:
                        raise RuntimeError("%s: cannot parse block member %r" % (name, mem))
                    if len(members) != 1:
                        raise RuntimeError("%s: block %s without instance name has %d members" % (name, bname, len(members)))
                    ty, nm, dim = m2.groups()
                    if dim is None:
                        out.append("%s& %s = *( %s* )( %s );" % (ty, nm, ty, src))
                    elif dim == "":
                        out.append("%s* const %s = ( %s* )( %s );" % (ty, nm, ty, src))
                    else:
                        out.append("%s ( &%s )[ %s ] = *( %s ( * )[ %s ] )( %s );" % (ty, nm, dim, ty, dim, src))
            pos = j + mm.end()
            continue
        raise RuntimeError("%s: unhandled layout declaration: layout(%s)%s" % (name, quals, rest[:120]))
    out.append(t[pos:])
    t = "".join(out)
    if info["local"] is None:
        raise RuntimeError(name + ": no local size")

    # GLM declares these two as well and argument-dependent lookup would find both: name the shim's (pinned rounding) explicitly
    t = re.sub(r"(?<![\w:])(packHalf2x16|unpackHalf2x16)\s*\(", r"glslshim::\1(", t)
    t = re.sub(r"\bshared\s+", "inline static thread_local ", t)
    # parameter qualifiers (the text has no comments any more)
    t = re.sub(r"([(,]\s*)(?:const\s+)?(?:inout|out)\s+(\w+)\s+(\w+)", r"\1\2& \3", t)
    t = re.sub(r"([(,]\s*)(const\s+)?in\s+(\w+)\s+(\w+)", r"\1\2\3 \4", t)
    t = FLOAT_LITERAL.sub(lambda m: m.group(1) + "f", t)

    # file-scope constants -> static constexpr members, unless a specialisation constant feeds them
    chunks = split_top_level(t)
    out = []
    for c in chunks:
        s = c.lstrip()
        if re.match(r"(?:const\s+)?\w+\s+\w+\s*\([^;{}=]*\)\s*;\s*$", s):
            continue  # a function prototype: members of a class need no forward declaration (and may not be redeclared)
        mm = re.match(r"const\s+(int|uint|float|bool)\s+\w+(\s*\[[^\]]*\])*\s*=", s)
        if mm and not any(re.search(r"\b%s\b" % n, s) for n in spec_names) and "glslshim::specConstant" not in s:
            c = c.replace("const", "static constexpr", 1)
        out.append(c)
    t = "".join(out)
    return t, info


def emit(name, body, info):
    ident = re.sub(r"\W", "_", os.path.splitext(name)[0])
    lx, ly, lz = info["local"]
    return ident, """// GENERATED by oracle/ref_glsl/build_ref.py from /root/reference/Source/Shaders/%s -- do not edit, do not commit
#include "glsl_shim.h"
namespace glslshim { namespace ref_%s // inside glslshim: its built-ins hide libm's ::exp2(double), ::abs(int), ...
{
struct Shader : glslshim::InvocationBase
{
%s
};
static_assert( sizeof( Shader::ShGlobalUniform ) == 2096, "ShGlobalUniform layout (Source/Generated/ShaderCommonC.h:209-304)" );
static void invoke() { Shader s; s.main(); }
} } // namespaces
extern "C" void asvgfRef_%s( const glslshim::DispatchContext* ctx, uint32_t groupsX, uint32_t groupsY )
{
    glslshim::dispatch( *ctx, glslshim::ref_%s::invoke, ( uint32_t )( %s ), ( uint32_t )( %s ), groupsX, groupsY, %s );
}
extern "C" void asvgfRefLocalSize_%s( uint32_t* xy ) { xy[ 0 ] = ( uint32_t )( %s ); xy[ 1 ] = ( uint32_t )( %s ); }
""" % (name, ident, body, ident, ident, lx, ly, "true" if info["barrier"] else "false", ident, lx, ly)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default=None)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    args = ap.parse_args()
    shaders_dir = os.path.join(args.reference, "Source", "Shaders")
    gen_dir = os.path.join(args.reference, "Source", "Generated")
    glm_dir = os.path.join(args.reference, "Tests", "Libs", "glm")
    if not os.path.isdir(shaders_dir):
        print("build_ref: %s not present, nothing to do" % shaders_dir)
        return 0
    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    cxx = ["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
           "-Werror=float-conversion",  # GLSL has no implicit float -> int conversion: none may sneak in through the shim
           "-I", HERE, "-isystem", glm_dir]
    objs, built = [], []
    procs = []
    for sh in SHADERS:
        if args.only and args.only not in sh:
            continue
        text = flatten(os.path.join(shaders_dir, sh), [shaders_dir, gen_dir], set())
        body, info = transform(preprocess(text), sh)
        ident, cpp = emit(sh, body, info)
        src = os.path.join(OUT, "gen", ident + ".cpp")
        open(src, "w").write(cpp)
        obj = os.path.join(OUT, "gen", ident + ".o")
        procs.append((sh, ident, subprocess.Popen(cxx + ["-c", src, "-o", obj], stderr=subprocess.PIPE, text=True)))
        objs.append(obj)
        if len(procs) >= args.jobs:
            pass
    failed = []
    for sh, ident, p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            failed.append(sh)
            sys.stderr.write("---- %s failed to compile ----\n%s\n" % (sh, "\n".join(err.splitlines()[:40])))
        else:
            built.append(ident)
    if failed:
        return 1
    shim_obj = os.path.join(OUT, "gen", "glsl_shim.o")
    subprocess.run(cxx + ["-c", os.path.join(HERE, "glsl_shim.cpp"), "-o", shim_obj], check=True)
    # the host-side LPM setup of ImageComposition::SetupLpmParams: AMD's headers compiled in their CPU mode where they lie
    lpm_obj = os.path.join(OUT, "gen", "lpm_setup_ref.o")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", os.path.join(shaders_dir, "LPM"),
                    "-c", os.path.join(HERE, "lpm_setup_ref.cpp"), "-o", lpm_obj], check=True)
    objs.append(lpm_obj)
    lib = os.path.join(OUT, "libasvgf_ref.so")
    subprocess.run(cxx + ["-shared", "-o", lib] + objs + [shim_obj], check=True)
    open(os.path.join(OUT, "shaders.txt"), "w").write("\n".join(built) + "\n")
    print("build_ref: %s (%d shaders: %s)" % (lib, len(built), ", ".join(built)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
