"""TEST INFRASTRUCTURE: turn pysph_b200/csrc/b200sph.cu into a translation unit g++ can
compile against tests/cpu_emul/cuda_shim.h.

Only three textual changes are made, everything else is the product source verbatim:
  * `#include <cuda_runtime.h>`            -> `#include "cuda_shim.h"`
  * the three inline-PTX helpers (frcp, frsqrt, ld_256) are dropped (the shim defines them)
  * `kernel<T..><<<grid, block, smem, stream>>>(args);`
                                            -> `emu::launch(grid, block, MODE, [&] { kernel<T..>(args); });`
    MODE is chosen per kernel from its own source: BLOCK if it uses __syncthreads / __shared__,
    WARP if it only uses warp shuffles / ballots / __syncwarp, SEQ otherwise.
"""
import re


def _match(text, i, open_ch, close_ch):
    """index just after the bracket that closes text[i] == open_ch"""
    depth = 0
    for k in range(i, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError('unbalanced %r at %d' % (open_ch, i))


def kernel_modes(src):
    modes = {}
    for m in re.finditer(r'__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(', src):
        name = m.group(1)
        j = _match(src, src.index('(', m.end() - 1), '(', ')')
        k = src.index('{', j) if src[j:j + 200].lstrip().startswith('{') else None
        if k is None:
            continue                              # a forward declaration
        body = src[k:_match(src, k, '{', '}')]
        called = [n for n in ('stage_body', 'stage_tvf_body', 'stage_solid_body', 'pair_body')
                  if n in body]
        if '__syncthreads' in body or '__shared__' in body:
            mode = 'emu::BLOCK'
        elif re.search(r'__shfl|__ballot|__syncwarp', body):
            mode = 'emu::WARP'
        else:
            mode = 'emu::SEQ'
        modes[name] = mode
    return modes


def transform(src):
    out = src.replace('#include <cuda_runtime.h>', '#include "cuda_shim.h"')
    for name in ('frcp', 'frsqrt', 'ld_256'):
        m = re.search(r'__device__ __forceinline__ \w+ %s\(' % name, out)
        k = out.index('{', m.start())
        out = out[:m.start()] + '// (%s: inline PTX in the CUDA build, defined by cuda_shim.h here)\n' % name + \
            out[_match(out, k, '{', '}'):]
    modes = kernel_modes(out)
    res, pos = [], 0
    for m in re.finditer(r'<<<', out):
        i = m.start()
        if i < pos:
            continue
        # kernel name (+ template arguments) before <<<
        j = i
        if out[j - 1] == '>':
            depth, j = 0, j - 1
            while True:
                if out[j] == '>':
                    depth += 1
                elif out[j] == '<':
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        k = j
        while out[k - 1].isalnum() or out[k - 1] == '_':
            k -= 1
        callee = out[k:i]
        name = re.match(r'\w+', callee).group(0)
        e = out.index('>>>', i)
        cfg = out[i + 3:e]
        parts, depth, cur = [], 0, ''
        for ch in cfg:
            if ch in '([':
                depth += 1
            elif ch in ')]':
                depth -= 1
            if ch == ',' and depth == 0:
                parts.append(cur)
                cur = ''
            else:
                cur += ch
        parts.append(cur)
        a0 = out.index('(', e)
        a1 = _match(out, a0, '(', ')')
        args = out[a0:a1]
        res.append(out[pos:k])
        res.append('emu::launch(%s, %s, %s, [&] { %s%s; })' % (parts[0].strip(), parts[1].strip(),
                                                                modes[name], callee, args))
        pos = a1
    res.append(out[pos:])
    return ''.join(res), modes


if __name__ == '__main__':
    import sys
    text, modes = transform(open(sys.argv[1]).read())
    open(sys.argv[2], 'w').write(text)
    for k in sorted(modes):
        print('%-24s %s' % (k, modes[k]))
