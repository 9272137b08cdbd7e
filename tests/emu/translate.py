"""Rewrites the CUDA launch statements  kernel<<<grid, block, smem, stream>>>(args);  of a .cu file into
emu::launch(dim3(grid), dim3(block), [&] { kernel(args); });  so that the library's host code compiles
with g++ for the CPU emulation of the tests.  Everything else is left to the preprocessor
(ONEPROT_KERNEL_EMULATION / ONEPROT_HOST_EMULATION).  Test infrastructure only.

    python tests/emu/translate.py in.cu out.cpp
"""
import re
import sys

LAUNCH = re.compile(r"((?:op|oph)::\w+(?:<[^<>;()]*>)?)\s*<<<")


def _match(text, i, open_ch, close_ch):
    depth = 0
    while True:
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1


def _split_top(s):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "(<":
            depth += 1
        elif c in ")>":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


def translate(text):
    out, pos = "", 0
    while True:
        m = LAUNCH.search(text, pos)
        if not m:
            return out + text[pos:]
        cfg_end = text.index(">>>", m.end())
        cfg = _split_top(text[m.end():cfg_end].replace("\\\n", " "))
        a0 = text.index("(", cfg_end)
        a1 = _match(text, a0, "(", ")")
        args = text[a0 + 1:a1]
        semi = a1 + 1
        assert text[semi] == ";", text[m.start():semi + 20]
        out += text[pos:m.start()] + f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&] {{ {m.group(1)}({args}); }});"
        pos = semi + 1


if __name__ == "__main__":
    open(sys.argv[2], "w").write(translate(open(sys.argv[1]).read()))
