#!/usr/bin/env python
"""pipes3.cu: issue cost of the integer / select / convert / shared-memory instruction kinds the merge loop can choose from
(same construction as gen_pipes2.py: one asm block per loop body; results stay live through dependent use)."""
import re
src = open("gen_pipes2.py").read()
head = src[:src.index("def k_ffma():")]
tail = src[src.index("DECL = "):]
kinds = '''
def k_ffma():
    return [f"fma.rn.f32 a{i}, b{i%4}, b{(i+1)%4}, a{i};" for i in range(16)] * 2
def k_ffma_imm():
    return [f"fma.rn.f32 a{i}, a{i}, 0f3F7FF000, b{i%4};" for i in range(16)] * 2
def k_iadd():
    return [f"add.s32 x{i}, x{(i+1)%16}, y{i%4};" for i in range(16)] * 2
def k_iadd3():
    return [f"add.s32 x{i}, x{(i+1)%16}, y{i%4}; add.s32 x{i}, x{i}, y{(i+2)%4};" for i in range(16)]
def k_isetp():
    return [f"setp.lt.and.s32 r{i%4}, x{i}, y{i%4}, r{i%4};" for i in range(16)] * 2
def k_mov():
    return [f"xor.b32 x{i}, x{i}, 1; mov.b32 x{(i+1)%16}, x{i};" for i in range(16)]
def k_prmt():
    return [f"prmt.b32 x{i}, x{i}, y{i%4}, 0x5410;" for i in range(16)] * 2
def k_sel():
    return [f"selp.b32 x{i}, x{(i+1)%16}, y{i%4}, pp{i%2};" for i in range(16)] * 2
def k_fmnmx():
    return [f"max.f32 a{i}, a{i}, b{i%4};" for i in range(16)] * 2
def k_i2f():
    return [f"cvt.rn.f32.s32 a{i}, x{i}; add.f32 a{i}, a{i}, b{i%4}; mov.b32 x{i}, a{i};" for i in range(16)]
def k_f2i():
    return [f"cvt.rzi.s32.f32 x{i}, a{i}; add.s32 x{i}, x{i}, y{i%4}; mov.b32 a{i}, x{i};" for i in range(16)]
def k_lop3p():
    return [f"and.b32 x{i}, x{(i+1)%16}, y{i%4}; setp.ne.s32 r{i%4}, x{i}, 0;" for i in range(16)]
def k_piadd():
    return [f"@pp{i%2} add.s32 x{i}, x{i}, y{i%4};" for i in range(16)] * 2
def k_pfadd():
    return [f"@pp{i%2} add.f32 a{i}, a{i}, b{i%4};" for i in range(16)] * 2
def k_bfe():
    return [f"bfe.s32 x{i}, x{(i+1)%16}, 3, 8;" for i in range(16)] * 2
def k_shl():
    return [f"shl.b32 x{i}, x{i}, 1; add.s32 x{i}, x{i}, y{i%4};" for i in range(16)]
def k_lea():
    return [f"shl.b32 x{i}, x{(i+1)%16}, 2; add.s32 x{i}, x{i}, y{i%4};" for i in range(16)]
def k_hfma2():
    return [f"fma.rn.f16x2 x{i}, y{i%4}, y{(i+1)%4}, x{i};" for i in range(16)] * 2
def k_cvt_f16():
    return [f"cvt.f32.f16 a{i}, h{i%4}; add.f32 a{i}, a{i}, b{i%4}; cvt.rn.f16.f32 h{i%4}, a{i};" for i in range(16)]
def k_vote():
    return [f"vote.sync.ballot.b32 x{i}, pp{i%2}, 0xffffffff;" for i in range(16)] * 2
def k_shfl():
    return [f"shfl.sync.bfly.b32 x{i}, x{i}, 1, 0x1f, 0xffffffff;" for i in range(16)] * 2
def k_lds32():
    return [f"ld.volatile.shared.f32 a{i}, [sa+{128*i}];" for i in range(16)] * 2
def k_lds64():
    return [f"ld.volatile.shared.v2.f32 {{a{2*i}, a{2*i+1}}}, [sb+{256*i}];" for i in range(8)] * 4
def k_lds128():
    return [f"ld.volatile.shared.v4.f32 {{a{4*i}, a{4*i+1}, a{4*i+2}, a{4*i+3}}}, [sc+{512*i}];" for i in range(4)] * 8
def k_lds32_2way():
    return [f"ld.volatile.shared.f32 a{i}, [sb+{256*i}];" for i in range(16)] * 2
def k_lds_s16():
    return [f"ld.volatile.shared.s16 x{i}, [sa+{128*i}];" for i in range(16)] * 2
def k_sts32():
    return [f"st.volatile.shared.f32 [sa+{128*i}], a{i};" for i in range(16)] * 2
def k_lds_ffma_1_3():
    out = []
    for i in range(8):
        out.append(f"ld.volatile.shared.f32 a{i}, [sa+{128*i}];")
        out += [f"fma.rn.f32 a{8+j}, b{j%4}, b{(j+1)%4}, a{8+j};" for j in range(3)]
    return out
def k_lds_ffma_1_1():
    out = []
    for i in range(16):
        out.append(f"ld.volatile.shared.f32 a{i%8}, [sa+{128*(i%8)}];")
        out.append(f"fma.rn.f32 a{8+i%8}, b{i%4}, b{(i+1)%4}, a{8+i%8};")
    return out
def k_lop3_ffma_1_3():
    out = []
    for i in range(8):
        out.append(f"lop3.b32 x{i}, x{i}, y{i%4}, y{(i+1)%4}, 0x96;")
        out += [f"fma.rn.f32 a{(3*i+j)%16}, b{j%4}, b{(j+1)%4}, a{(3*i+j)%16};" for j in range(3)]
    return out
def k_fsel_ffma_1_3():
    out = []
    for i in range(8):
        out.append(f"selp.f32 a{i}, a{(i+1)%8}, b{i%4}, pp{i%2};")
        out += [f"fma.rn.f32 a{8+(3*i+j)%8}, b{j%4}, b{(j+1)%4}, a{8+(3*i+j)%8};" for j in range(3)]
    return out

TESTS = [("FFMA", k_ffma), ("FFMA imm", k_ffma_imm), ("IADD", k_iadd), ("IADD3 (2 adds)", k_iadd3), ("ISETP.and", k_isetp), ("MOV", k_mov), ("PRMT", k_prmt),
         ("SEL", k_sel), ("FMNMX", k_fmnmx), ("I2F", k_i2f), ("F2I", k_f2i), ("AND+SETP", k_lop3p), ("@p IADD", k_piadd), ("@p FADD", k_pfadd),
         ("BFE", k_bfe), ("SHL", k_shl), ("SHL+ADD", k_lea), ("HFMA2", k_hfma2), ("CVT f16->f32", k_cvt_f16), ("VOTE", k_vote), ("SHFL", k_shfl),
         ("LDS.32", k_lds32), ("LDS.64", k_lds64), ("LDS.128", k_lds128), ("LDS.32 2-way conflict", k_lds32_2way), ("LDS.S16", k_lds_s16), ("STS.32", k_sts32),
         ("LDS+3 FFMA", k_lds_ffma_1_3), ("LDS+FFMA 1:1", k_lds_ffma_1_1), ("LOP3+3 FFMA", k_lop3_ffma_1_3), ("FSEL+3 FFMA", k_fsel_ffma_1_3)]
'''
exec_src = head + kinds + tail.replace('open("pipes2.cu", "w")', 'open("pipes3.cu", "w")')
exec_src = re.sub(r'TESTS = \[\("FFMA", k_ffma\), \("FMUL".*?\n\s+\("LDS\.32", k_lds32\).*?\]\n', '', exec_src, flags=re.S)
open("_gen3_tmp.py", "w").write(exec_src)
