#!/usr/bin/env python3
"""R&D harness (not part of the product): issue-slot / pipe-sharing microbenchmarks for sm_100a.

Generates one PTX kernel per variant (loop bodies made of FFMA2 / FFMA / FSETP / predicated adds /
LEA.HI ... in chosen ratios, all operands loop-varying so that ptxas cannot hoist anything),
assembles each with `ptxas -arch=sm_100a` to tools/ptx_lab/out/<name>.cubin and prints the SASS opcode
histogram of the loop so that the instruction mix is verified HERE before GPU time is spent.
tools/ptx_lab/run.cu loads the cubins on the GPU box and times them.

    python tools/ptx_lab/gen.py            # writes out/*.ptx, out/*.cubin, out/manifest.txt
"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "out")

HEADER = """.version 8.8
.target sm_100a
.address_size 64
.const .align 16 .b8 cpts[49152];
.visible .entry k(.param .u32 p_iters, .param .u64 p_sink, .param .f32 p_seed)
.maxntid 256, 1, 1
{
  .reg .b64 A<16>, B<8>, C<16>, W<16>, sink;
  .reg .f32 fa<16>, fb<8>, fc<16>, lo<16>, hi<16>, wl<16>, wh<16>, thr, nthr2, seed, ftid, tmpf, x<8>;
  .reg .pred p<8>, q<8>, ploop;
  .reg .u32 cnt<16>, t<16>, iters, it, tid, one, msk;
  .reg .b32 h<16>;
  .reg .b64 coff, cbase, ck0, ck1, gt0, gt1, PX<2>, PY<2>, PZ<2>, T<16>, HD<16>, kneg;
  .reg .f32 hn<48>, sl<16>, sh<16>;
  .reg .u32 saddr, soff;
  .shared .align 16 .b8 tile[6144];
  .reg .u32 ctaid;
  .reg .pred pfirst;
  ld.param.u32 iters, [p_iters];
  ld.param.u64 sink, [p_sink];
  ld.param.f32 seed, [p_seed];
  mov.u32 tid, %tid.x;
  cvt.rn.f32.u32 ftid, tid;
  fma.rn.f32 seed, ftid, 0f33D6BF95, seed;      // + tid * 1e-7
  fma.rn.f32 thr, seed, 0f2F000000, 0f3F000000;     // ~0.5, run-time value
  fma.rn.f32 nthr2, seed, 0f2F000000, 0fBE800000;   // ~-0.25
  mov.u32 one, 1;
  mov.b32 msk, thr;
  neg.s32 msk, msk;
  cvt.u64.u32 kneg, msk;                            // 2^32 - bits(thr): the addend of the carry-chain count
  // fill the shared tile (3 components x 512 floats) with run-time values
  mov.u32 saddr, tile;
  shl.b32 soff, tid, 2;
  add.u32 soff, soff, saddr;
  st.shared.f32 [soff], seed;
  st.shared.f32 [soff+1024], seed;
  st.shared.f32 [soff+2048], seed;
  st.shared.f32 [soff+3072], seed;
  st.shared.f32 [soff+4096], seed;
  st.shared.f32 [soff+5120], seed;
  bar.sync 0;
  mov.u64 ck0, %clock64;
  mov.u64 gt0, %globaltimer;
"""


def f32hex(v):
    import struct
    return "0f%08X" % struct.unpack("<I", struct.pack("<f", v))[0]


def init():
    s = []
    for i in range(16):
        s.append(f"  add.f32 lo{i}, seed, {f32hex(0.001 * i)};")
        s.append(f"  sub.f32 hi{i}, seed, {f32hex(0.002 * i)};")
        s.append(f"  mov.b64 A{i}, {{lo{i}, hi{i}}};")
        s.append(f"  mov.f32 fa{i}, lo{i};")
        s.append(f"  fma.rn.f32 fc{i}, seed, {f32hex(1e-6 * (i + 1))}, {f32hex(1.37e-7 * (i + 3))};")
        s.append(f"  fma.rn.f32 tmpf, seed, {f32hex(1.7e-6 * (i + 1))}, {f32hex(2.11e-7 * (i + 5))};")
        s.append(f"  mov.b64 C{i}, {{fc{i}, tmpf}};")
        s.append(f"  mov.u32 cnt{i}, {i};")
    for i in range(48):
        s.append(f"  fma.rn.f32 hn{i}, seed, {f32hex(1.1e-4 * (i + 1))}, {f32hex(0.01 * (i + 1))};")
    for i in range(8):
        s.append(f"  fma.rn.f32 fb{i}, seed, {f32hex(1e-9 * (i + 1))}, {f32hex(0.99999 - 1e-6 * i)};")
        s.append(f"  fma.rn.f32 tmpf, seed, {f32hex(1.3e-9 * (i + 2))}, {f32hex(0.99998 + 1.1e-6 * i)};")
        s.append(f"  mov.b64 B{i}, {{fb{i}, tmpf}};")
        s.append(f"  setp.lt.f32 p{i}, seed, {f32hex(100.0)};")
    return "\n".join(s) + "\n"


FOOTER_TMPL = """
  add.u32 it, it, 1;
  setp.lt.u32 ploop, it, iters;
  @ploop bra LOOP;
  mov.u64 ck1, %clock64;
  mov.u64 gt1, %globaltimer;
  mov.u32 ctaid, %ctaid.x;
  or.b32 ctaid, ctaid, tid;
  setp.eq.u32 pfirst, ctaid, 0;
  sub.u64 ck1, ck1, ck0;
  sub.u64 gt1, gt1, gt0;
  @pfirst st.global.u64 [sink+8], ck1;
  @pfirst st.global.u64 [sink+16], gt1;
{consume}
  setp.eq.u32 ploop, cnt0, 0x12345678;
  @ploop st.global.u32 [sink], cnt0;
  ret;
}}
"""


def consume():
    s = []
    for i in range(16):
        s.append(f"  mov.b64 {{lo{i}, hi{i}}}, A{i};")
        s.append(f"  mov.b32 t0, lo{i}; add.u32 cnt0, cnt0, t0; mov.b32 t0, hi{i}; add.u32 cnt0, cnt0, t0;")
        s.append(f"  mov.b32 t0, fa{i}; add.u32 cnt0, cnt0, t0;")
        s.append(f"  mov.b32 t0, fc{i}; add.u32 cnt0, cnt0, t0;")
        if i:
            s.append(f"  add.u32 cnt0, cnt0, cnt{i};")
    for i in range(8):
        s.append(f"  selp.u32 t0, 1, 0, p{i}; add.u32 cnt0, cnt0, t0;")
    for i in range(16):
        s.append(f"  mov.b64 {{sl{i}, sh{i}}}, T{i}; mov.b32 t0, sl{i}; add.u32 cnt0, cnt0, t0; mov.b32 t0, sh{i}; add.u32 cnt0, cnt0, t0;")
        s.append(f"  mov.b64 {{t0, t1}}, HD{i}; add.u32 cnt0, cnt0, t1;")
    return "\n".join(s)


class Body:
    """Accumulates PTX lines for one unrolled slice; u = slice number (rotates operands)."""

    def __init__(self):
        self.lines = []
        self.unpacked = set()
        self.nsrc = 16          # accumulators that the main instructions of the slice keep loop-varying

    def emit(self, s):
        self.lines.append("  " + s)

    # ---- instruction kinds -------------------------------------------------------------
    def ffma2(self, i, u):
        self.emit(f"fma.rn.f32x2 A{i % 16}, A{i % 16}, B{i % 8}, C{i % 16};")
        self.unpacked.discard(i % 16)

    def ffma2_bcast(self, i, u):  # one scalar operand broadcast to both halves (.F32 operand form)
        self.emit(f"mov.b64 W15, {{fb{i % 8}, fb{i % 8}}};")
        self.emit(f"fma.rn.f32x2 A{i % 16}, W15, A{i % 16}, C{i % 16};")
        self.unpacked.discard(i % 16)

    def ffma(self, i, u):
        self.emit(f"fma.rn.f32 fa{i % 16}, fa{i % 16}, fb{i % 8}, fc{(i + u) % 16};")

    def rf_a(self, i, u):
        self.emit(f"fma.rn.f32x2 A{i % 16}, A{i % 16}, B0, C0;")

    def rf_b(self, i, u):
        self.emit(f"fma.rn.f32x2 A{i % 16}, A{i % 16}, B{i % 8}, C0;")

    def rf_c(self, i, u):
        self.emit(f"mov.b64 W15, {{fb{i % 8}, fb{i % 8}}};")
        self.emit(f"fma.rn.f32x2 A{i % 16}, W15, B0, A{i % 16};")

    def rf_d(self, i, u):
        self.emit(f"mov.b64 W15, {{fb{i % 8}, fb{i % 8}}};")
        self.emit(f"mov.b64 W14, {{fc{i % 16}, fc{i % 16}}};")
        self.emit(f"fma.rn.f32x2 A{i % 16}, W15, A{i % 16}, W14;")

    def rf_e(self, i, u):
        self.emit(f"fma.rn.f32 fa{i % 16}, fa{i % 16}, fb0, fc{i % 16};")

    def rf_f(self, i, u):
        self.emit(f"fma.rn.f32 fa{i % 16}, fa{i % 16}, fb0, fc0;")

    def rf_g(self, i, u):  # 2 pair operands only: A = A*A + C
        self.emit(f"fma.rn.f32x2 A{i % 16}, A{i % 16}, A{i % 16}, C{i % 16};")

    def unpack(self, i):
        i %= 16
        if i not in self.unpacked:
            self.emit(f"mov.b64 {{lo{i}, hi{i}}}, A{i};")
            self.unpacked.add(i)

    def src(self, j, scalar):
        """j-th loop-varying float source: halves of the FFMA2 accumulators, or the scalar FFMA accumulators."""
        if scalar:
            return f"fa{j % self.nsrc}"
        i = (j // 2) % self.nsrc
        self.unpack(i)
        return (f"lo{i}" if j % 2 == 0 else f"hi{i}")

    def setp_chain(self, j, scalar):  # FSETP with a predicate input: p = (|x| < thr) & p
        x = self.src(j, scalar)
        self.emit(f"abs.f32 x0, {x};")
        self.emit(f"setp.lt.and.f32 p{j % 4}, x0, thr, p{j % 4};")

    def setp_padd(self, j, scalar):  # FSETP + @p add
        x = self.src(j, scalar)
        self.emit(f"abs.f32 x0, {x};")
        self.emit(f"setp.lt.f32 q0, x0, thr;")
        self.emit(f"@q0 add.u32 cnt{j % 8}, cnt{j % 8}, 1;")

    def setp_padd4(self, j, scalar):  # 4 FSETP then 4 predicated adds into one counter (as the product kernel does)
        xs = [self.src(4 * j + k, scalar) for k in range(4)]
        for k in range(4):
            self.emit(f"abs.f32 x{k}, {xs[k]};")
        for k in range(4):
            self.emit(f"setp.lt.f32 q{k}, x{k}, thr;")
        for k in range(4):
            self.emit(f"@q{k} add.u32 cnt{j % 8}, cnt{j % 8}, 1;")

    def setp2_selp_add(self, j, scalar):  # 2 FSETP + (selp, selp, add, add): does ptxas find a 2-predicate add?
        xs = [self.src(2 * j + k, scalar) for k in range(2)]
        for k in range(2):
            self.emit(f"abs.f32 x{k}, {xs[k]};")
            self.emit(f"setp.lt.f32 q{k}, x{k}, thr;")
        self.emit("selp.u32 t0, 1, 0, q0;")
        self.emit("selp.u32 t1, 1, 0, q1;")
        self.emit(f"add.u32 t0, t0, t1;")
        self.emit(f"add.u32 cnt{j % 8}, cnt{j % 8}, t0;")

    def setp2_addc(self, j, scalar):
        """2 FSETP + one 3-input add with carry-style predicates: cnt = cnt + q0 + q1 written as
        two selp-negated masks subtracted (IADD3 cnt, -m0, -m1)."""
        xs = [self.src(2 * j + k, scalar) for k in range(2)]
        for k in range(2):
            self.emit(f"abs.f32 x{k}, {xs[k]};")
        self.emit(f"set.lt.u32.f32 t0, x0, thr;")
        self.emit(f"set.lt.u32.f32 t1, x1, thr;")
        self.emit(f"sub.u32 cnt{j % 8}, cnt{j % 8}, t0;")
        self.emit(f"sub.u32 cnt{j % 8}, cnt{j % 8}, t1;")

    def fset_fadd(self, j, scalar):  # set.lt.f32.f32 (1.0f / 0) + float accumulate
        x = self.src(j, scalar)
        self.emit(f"abs.f32 x0, {x};")
        self.emit(f"set.lt.f32.f32 x1, x0, thr;")
        self.emit(f"add.f32 fc{8 + j % 8}, fc{8 + j % 8}, x1;")

    def min3_setp(self, j, scalar):  # 4 values -> min3, min -> one FSETP chain
        xs = [self.src(4 * j + k, scalar) for k in range(4)]
        for k in range(4):
            self.emit(f"abs.f32 x{k}, {xs[k]};")
        self.emit("min.f32 x4, x0, x1, x2;")
        self.emit("min.f32 x4, x4, x3;")
        self.emit(f"setp.lt.and.f32 p{j % 4}, x4, thr, p{j % 4};")

    def fsetbf(self, j, scalar):   # FSET.BF only; results folded with a 2-input add every other op to keep them live
        x = self.src(j, scalar)
        self.emit(f"abs.f32 x0, {x};")
        self.emit(f"set.lt.f32.f32 x1, x0, thr;")
        self.emit(f"mov.b32 h0, x1; add.u32 cnt{j % 8}, cnt{j % 8}, h0;")

    def raw3(self, j, scalar):     # the product kernel's counting: 2 FSET.BF + one 3-input add
        a, b = self.src(2 * j, scalar), self.src(2 * j + 1, scalar)
        self.emit(f"abs.f32 x0, {a}; abs.f32 x1, {b};")
        self.emit(f"set.lt.f32.f32 x2, x0, thr; set.lt.f32.f32 x3, x1, thr;")
        self.emit(f"mov.b32 h0, x2; mov.b32 h1, x3; add.u32 t0, h0, h1; add.u32 cnt{j % 8}, cnt{j % 8}, t0;")

    def leahi(self, j, scalar):
        x = self.src(j, scalar)
        self.emit(f"mov.b32 h0, {x};")
        self.emit("shr.u32 t0, h0, 31;")
        self.emit(f"add.u32 cnt{j % 8}, cnt{j % 8}, t0;")

    def imadhi(self, j, scalar):  # cnt += (x*2) >> 32  == x >> 31, on the FMA-heavy pipe
        x = self.src(j, scalar)
        self.emit(f"mov.b32 h0, {x};")
        self.emit(f"mad.hi.u32 cnt{j % 8}, h0, 2, cnt{j % 8};")

    def iadd3(self, j, scalar):
        a, b = self.src(2 * j, scalar), self.src(2 * j + 1, scalar)
        self.emit(f"mov.b32 h0, {a}; mov.b32 h1, {b};")
        self.emit(f"add.u32 t0, h0, h1;")
        self.emit(f"add.u32 cnt{j % 8}, cnt{j % 8}, t0;")

    def lop3(self, j, scalar):
        a, b = self.src(2 * j, scalar), self.src(2 * j + 1, scalar)
        self.emit(f"mov.b32 h0, {a}; mov.b32 h1, {b};")
        self.emit(f"lop3.b32 cnt{j % 8}, cnt{j % 8}, h0, h1, 0x96;")

    def prmt(self, j, scalar):
        a, b = self.src(2 * j, scalar), self.src(2 * j + 1, scalar)
        self.emit(f"mov.b32 h0, {a}; mov.b32 h1, {b};")
        self.emit(f"prmt.b32 h2, h0, h1, 0xbf3b;")  # sign-replicated top bytes
        self.emit(f"sub.u32 cnt{j % 8}, cnt{j % 8}, h2;")

    def f2fp(self, j, scalar):  # pack two floats to half2, keep sign bits, shift-add
        a, b = self.src(2 * j, scalar), self.src(2 * j + 1, scalar)
        self.emit(f"cvt.rn.f16x2.f32 h0, {a}, {b};")
        self.emit(f"and.b32 h0, h0, 0x80008000;")
        self.emit(f"shr.u32 h0, h0, 15;")
        self.emit(f"add.u32 cnt{j % 8}, cnt{j % 8}, h0;")

    def sq_sign2(self, i):  # w = a*a - thr^2 packed (extra FFMA2 of the sign form); result in W
        self.emit(f"mov.b64 W14, {{nthr2, nthr2}};")
        self.emit(f"fma.rn.f32x2 W{i % 14}, A{i % 16}, A{i % 16}, W14;")

    def leahi_w(self, i, j):
        self.emit(f"mov.b64 {{wl{i % 14}, wh{i % 14}}}, W{i % 14};")
        self.emit(f"mov.b32 h0, wl{i % 14}; shr.u32 t0, h0, 31; add.u32 cnt{j % 8}, cnt{j % 8}, t0;")
        self.emit(f"mov.b32 h1, wh{i % 14}; shr.u32 t1, h1, 31; add.u32 cnt{j % 8}, cnt{j % 8}, t1;")


def interleave(b, u, main, nmain, aux, naux, scalar=False):
    """Emit nmain main instructions and naux aux instructions, evenly interleaved."""
    done_aux = 0
    for i in range(nmain):
        main(i, u)
        want = (i + 1) * naux // max(nmain, 1)
        while done_aux < want:
            aux(done_aux + u * naux, scalar)
            done_aux += 1
    while done_aux < naux:
        aux(done_aux + u * naux, scalar)
        done_aux += 1


VARIANTS = collections.OrderedDict()


def variant(name, desc):
    def deco(fn):
        VARIANTS[name] = (desc, fn)
        return fn
    return deco


def simple(name, desc, main, nmain, aux=None, naux=0, scalar=False):
    def fn(b, u):
        m = getattr(b, main)
        b.nsrc = min(nmain, 16)
        if aux is None:
            for i in range(nmain):
                m(i, u)
        else:
            interleave(b, u, m, nmain, getattr(b, aux), naux, scalar)
    VARIANTS[name] = (desc, fn)


simple("f2_12", "12 FFMA2", "ffma2", 12)
simple("f2b_12", "12 FFMA2 (one .F32 broadcast operand)", "ffma2_bcast", 12)
simple("f1_12", "12 FFMA", "ffma", 12)
simple("rf_a", "12 FFMA2  A=A*Bsame+Csame", "rf_a", 12)
simple("rf_b", "12 FFMA2  A=A*B_i+Csame", "rf_b", 12)
simple("rf_c", "12 FFMA2  A=bcast(b_i)*Xsame+A", "rf_c", 12)
simple("rf_d", "12 FFMA2  A=bcast(b_i)*A+bcast(c_i)", "rf_d", 12)
simple("rf_e", "12 FFMA   a=a*bsame+c_i", "rf_e", 12)
simple("rf_f", "12 FFMA   a=a*bsame+csame", "rf_f", 12)
simple("rf_g", "12 FFMA2  A=A*A+C_i", "rf_g", 12)
for k in (4, 8, 12):
    simple(f"f2_12_setpc_{k}", f"12 FFMA2 + {k} FSETP(pred chain)", "ffma2", 12, "setp_chain", k)
    simple(f"f2_12_leahi_{k}", f"12 FFMA2 + {k} LEA.HI", "ffma2", 12, "leahi", k)
simple("f2_12_imadhi_8", "12 FFMA2 + 8 IMAD.HI (cnt += x>>31 on the FMA-heavy pipe)", "ffma2", 12, "imadhi", 8)
simple("f2_12_iadd3_8", "12 FFMA2 + 8 IADD3", "ffma2", 12, "iadd3", 8)
simple("f2_12_lop3_8", "12 FFMA2 + 8 LOP3", "ffma2", 12, "lop3", 8)
simple("f2_12_sp_8", "12 FFMA2 + 8 (FSETP + @p add)", "ffma2", 12, "setp_padd", 8)
simple("f2_12_sp4_2", "12 FFMA2 + 2 x (4 FSETP + 4 @p add)  [product ratio 3:2:2]", "ffma2", 12, "setp_padd4", 2)
simple("f2_12_s2sel_4", "12 FFMA2 + 4 x (2 FSETP + selp,selp,add,add)", "ffma2", 12, "setp2_selp_add", 4)
simple("f2_12_s2set_4", "12 FFMA2 + 4 x (2 set.lt.u32 + sub,sub)", "ffma2", 12, "setp2_addc", 4)
simple("f2_12_prmt_4", "12 FFMA2 + 4 x (PRMT sign bytes + sub)", "ffma2", 12, "prmt", 4)
simple("f2_12_f2fp_4", "12 FFMA2 + 4 x (F2FP pack + and + shr + add)", "ffma2", 12, "f2fp", 4)
simple("f2_12_fset_8", "12 FFMA2 + 8 x (set.lt.f32.f32 + FADD)", "ffma2", 12, "fset_fadd", 8)
simple("f2_12_min3_2", "12 FFMA2 + 2 x (min3, min, FSETP chain) over 4 values", "ffma2", 12, "min3_setp", 2)
simple("f1_12_setpc_4", "12 FFMA + 4 FSETP(pred chain)", "ffma", 12, "setp_chain", 4, True)
simple("f1_12_sp_4", "12 FFMA + 4 (FSETP + @p add)", "ffma", 12, "setp_padd", 4, True)
simple("f1_12_leahi_4", "12 FFMA + 4 LEA.HI", "ffma", 12, "leahi", 4, True)
simple("f1_8_fsetbf_8", "8 FFMA + 8 x (FSET.BF + 2-input add)", "ffma", 8, "fsetbf", 8, True)
simple("f1_8_raw3_4", "8 FFMA + 4 x (2 FSET.BF + IADD3 with 3 register inputs)", "ffma", 8, "raw3", 4, True)
simple("f1_8_iadd3_4", "8 FFMA + 4 IADD3 (3 register inputs)", "ffma", 8, "iadd3", 4, True)
simple("f1_8_leahi_8", "8 FFMA + 8 LEA.HI", "ffma", 8, "leahi", 8, True)
simple("f1_8_setpc_8", "8 FFMA + 8 FSETP (pred chain)", "ffma", 8, "setp_chain", 8, True)
simple("f2_8_raw3_4", "8 FFMA2 + 4 x (2 FSET.BF + IADD3)  [2D line ratio 2:2:1]", "ffma2", 8, "raw3", 4)
simple("f2_12_raw3_4", "12 FFMA2 + 4 x (2 FSET.BF + IADD3)  [plane ratio 3:2:1]", "ffma2", 12, "raw3", 4)
simple("f2_16_raw3_4", "16 FFMA2 + 4 x (2 FSET.BF + IADD3)  [4:2:1]", "ffma2", 16, "raw3", 4)
simple("setpc_8", "8 FSETP(pred chain) only (sources: scalar FFMA accumulators, 2 FFMA to keep them varying)", "ffma", 2, "setp_chain", 8, True)
simple("leahi_8", "8 LEA.HI only (+2 FFMA)", "ffma", 2, "leahi", 8, True)
simple("sp_8", "8 (FSETP + @p add) only (+2 FFMA)", "ffma", 2, "setp_padd", 8, True)


@variant("sign_16", "sign form: per slice 12 FFMA2 + 4 FFMA2(w=s*s-d2) + 8 LEA.HI  (= 4 pair-evals)")
def _sign(b, u):
    for g in range(4):
        for i in range(3):
            b.ffma2(3 * g + i, u)
        b.sq_sign2(3 * g + 2)
        b.leahi_w(3 * g + 2, g + u)


def plane(order, R, count, src="lds"):
    """One slice = 2 point pairs (one LDS.128 per component) x R hypotheses = 4R evals.
    order: 'hm' hypothesis-major (chain per hypothesis), 'cm' component-major (all R hypotheses per component).
    count: 'none' (min3 consumer), 'setp' (FSETP + @p add), 'sign' (w = s*s - d2, LEA.HI), 'setpc' (FSETP pred chain)"""
    def fn(b, u):
        if src == "lds":
            b.emit(f"shl.b32 soff, it, 6; add.u32 soff, soff, {16 * u}; and.b32 soff, soff, 2032; add.u32 soff, soff, saddr;")
            b.emit("ld.shared.v2.b64 {PX0, PX1}, [soff];")
            b.emit("ld.shared.v2.b64 {PY0, PY1}, [soff+2048];")
            b.emit("ld.shared.v2.b64 {PZ0, PZ1}, [soff+4096];")
        else:
            b.emit(f"shl.b32 soff, it, 6; add.u32 soff, soff, {16 * u}; and.b32 soff, soff, 16368; cvt.u64.u32 coff, soff;")
            b.emit("mov.u64 cbase, cpts; add.u64 cbase, cbase, coff;")
            b.emit("ld.const.v2.b64 {PX0, PX1}, [cbase];")
            b.emit("ld.const.v2.b64 {PY0, PY1}, [cbase+16384];")
            b.emit("ld.const.v2.b64 {PZ0, PZ1}, [cbase+32768];")
        def bc(name):
            return "{" + name + ", " + name + "}"
        for pr in range(2):
            X, Y, Z = f"PX{pr}", f"PY{pr}", f"PZ{pr}"
            if order == "hm":
                for r in range(R):
                    if count == "chain":
                        b.emit(f"mov.b64 W0, {bc(f'hn{4*r+2}')}; fma.rn.f32x2 T{r}, W0, {Z}, T{r};")
                    else:
                        b.emit(f"mov.b64 W0, {bc(f'hn{4*r+2}')}; mov.b64 W1, {bc(f'hn{4*r+3}')}; fma.rn.f32x2 T{r}, W0, {Z}, W1;")
                    b.emit(f"mov.b64 W0, {bc(f'hn{4*r+1}')}; fma.rn.f32x2 T{r}, W0, {Y}, T{r};")
                    b.emit(f"mov.b64 W0, {bc(f'hn{4*r}')}; fma.rn.f32x2 T{r}, W0, {X}, T{r};")
                    counting(b, r, count, pr)
            else:
                for r in range(R):
                    b.emit(f"mov.b64 W0, {bc(f'hn{4*r+2}')}; mov.b64 W1, {bc(f'hn{4*r+3}')}; fma.rn.f32x2 T{r}, W0, {Z}, W1;")
                for r in range(R):
                    b.emit(f"mov.b64 W0, {bc(f'hn{4*r+1}')}; fma.rn.f32x2 T{r}, W0, {Y}, T{r};")
                for r in range(R):
                    b.emit(f"mov.b64 W0, {bc(f'hn{4*r}')}; fma.rn.f32x2 T{r}, W0, {X}, T{r};")
                for r in range(R):
                    counting(b, r, count, pr)
    return fn


def counting(b, r, count, pr):
    if count == "chain":
        return
    b.emit(f"mov.b64 {{sl{r}, sh{r}}}, T{r};")
    if count == "raw":      # 2 FSET.BF + one 3-input add of the raw words (the product kernel's form)
        b.emit(f"abs.f32 x0, sl{r}; abs.f32 x1, sh{r}; set.lt.f32.f32 x2, x0, thr; set.lt.f32.f32 x3, x1, thr;")
        b.emit(f"mov.b32 h0, x2; mov.b32 h1, x3; add.u32 t0, h0, h1; add.u32 cnt{r}, cnt{r}, t0;")
        return
    if count == "carry":    # the product kernel's plane form: two carry-only IADD3 + one IADD3.X (64-bit add idiom, high word = counter)
        for half in ("sl", "sh"):
            b.emit(f"mov.b32 h0, {half}{r}; mov.b64 {{t0, t1}}, HD{r}; mov.b64 W2, {{h0, t1}}; add.u64 HD{r}, W2, kneg;")
        return
    if count == "fset":     # 2 FSET.BF, results consumed by one LOP3 (keeps the ALU count at 3 but with no carry chain)
        b.emit(f"abs.f32 x0, sl{r}; abs.f32 x1, sh{r}; set.lt.f32.f32 x2, x0, thr; set.lt.f32.f32 x3, x1, thr;")
        b.emit(f"mov.b32 h0, x2; mov.b32 h1, x3; lop3.b32 cnt{r}, cnt{r}, h0, h1, 0x96;")
        return
    if count == "fset1":    # one FSET.BF per pair only (half the compares): where does the time go?
        b.emit(f"abs.f32 x0, sl{r}; set.lt.f32.f32 x2, x0, thr; mov.b32 h0, x2; mov.b32 h1, sh{r}; add.u32 t0, h0, h1; add.u32 cnt{r}, cnt{r}, t0;")
        return
    if count == "setp":
        b.emit(f"abs.f32 x0, sl{r}; abs.f32 x1, sh{r}; setp.lt.f32 q0, x0, thr; setp.lt.f32 q1, x1, thr;")
        b.emit(f"@q0 add.u32 cnt{r}, cnt{r}, 1; @q1 add.u32 cnt{r}, cnt{r}, 1;")
    elif count == "setpc":
        b.emit(f"abs.f32 x0, sl{r}; abs.f32 x1, sh{r}; setp.lt.and.f32 p{r % 4}, x0, thr, p{r % 4}; setp.lt.and.f32 p{r % 4}, x1, thr, p{r % 4};")
    elif count == "sign":
        b.emit(f"mov.b64 W2, {{nthr2, nthr2}}; fma.rn.f32x2 W3, T{r}, T{r}, W2; mov.b64 {{wl0, wh0}}, W3;")
        b.emit(f"mov.b32 h0, wl0; shr.u32 t0, h0, 31; add.u32 cnt{r}, cnt{r}, t0; mov.b32 h1, wh0; shr.u32 t1, h1, 31; add.u32 cnt{r}, cnt{r}, t1;")
    elif count == "none":
        b.emit(f"abs.f32 x0, sl{r}; abs.f32 x1, sh{r}; min.f32 fc{8 + r % 8}, fc{8 + r % 8}, x0, x1;")


for R in (8,):
    for order in ("hm", "cm"):
        for count in ("none", "setpc", "setp", "sign"):
            VARIANTS[f"pl_{order}_{count}_R{R}"] = (f"plane kernel body, {order}, R={R}, counting={count}: {4*R} evals/slice, {6*R} FFMA2 (+{2*R} for sign)", plane(order, R, count))
for R in (8, 10):
    for count in ("chain", "raw", "carry", "fset", "fset1"):
        VARIANTS[f"plc_{count}_R{R}"] = (f"plane body, points from the constant bank (LDCU -> UR operands), R={R}, counting={count}: {4*R} evals/slice", plane("hm", R, count, "const"))
for R in (8, 12):
    for count in ("none", "setpc", "setp", "sign"):
        VARIANTS[f"plc_{count}_R{R}"] = (f"plane body, points from the constant bank (LDCU -> UR operands), R={R}, counting={count}: {4*R} evals/slice", plane("hm", R, count, "const"))


def build(name, fn, unroll=4):
    body = []
    for u in range(unroll):
        b = Body()
        fn(b, u)
        body += b.lines
    ptx = HEADER + init() + "  mov.u32 it, 0;\nLOOP:\n" + "\n".join(body) + FOOTER_TMPL.format(consume=consume())
    os.makedirs(OUT, exist_ok=True)
    pp = os.path.join(OUT, name + ".ptx")
    cb = os.path.join(OUT, name + ".cubin")
    open(pp, "w").write(ptx)
    subprocess.check_call(["ptxas", "-arch=sm_100a", "-O3", pp, "-o", cb])
    sass = subprocess.check_output(["cuobjdump", "-sass", cb], text=True)
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    # loop = from the branch target of the last backward BRA to that BRA
    loop = None
    for k, (addr, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr and loop is None:
            tgt = int(m.group(1), 16)
            loop = [t for a, t in ins if tgt <= a <= addr]
    hist = collections.Counter()
    hist["(.reuse flags)"] = sum(t.count(".reuse") for t in loop or [])
    for t in loop or []:
        t = re.sub(r"^@!?U?P\d\s+", "@p ", t)
        op = t.split()[0] if not t.startswith("@p") else "@p " + t.split()[1]
        hist[op] += 1
    return hist, len(loop or [])


def main():
    only = sys.argv[1:]
    man = []
    for name, (desc, fn) in VARIANTS.items():
        if only and name not in only:
            continue
        hist, n = build(name, fn)
        per = {k: v / 4 for k, v in hist.items()}
        hs = " ".join(f"{k}={v:g}" for k, v in sorted(per.items(), key=lambda kv: -kv[1]))
        print(f"{name:18s} {desc}\n    per slice ({n / 4:g} instr): {hs}")
        man.append(f"{name}\t{n / 4:g}\t{desc}\t{hs}")
    open(os.path.join(OUT, "manifest.txt"), "w").write("\n".join(man) + "\n")


if __name__ == "__main__":
    main()
