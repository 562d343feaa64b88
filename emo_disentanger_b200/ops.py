"""Thin tensor->pointer wrappers over the C ABI (include/emo_b200.h).  torch tensors are storage
only: every function here launches hand-written sm_100a kernels on torch's current stream."""
import ctypes as C
import torch

from . import _lib as L
from ._lib import F32, BF16, GEMM_NT, GEMM_NN, GEMM_TN, ACT_NONE, ACT_RELU, ACT_GELU_NEW, \
    ACT_RELU_MASK_BWD, ACT_GELU_NEW_BWD  # noqa: F401


def _p(t):
    return None if t is None else t.data_ptr()


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("emo ops take float32 or bfloat16 tensors, got %s" % t.dtype)


def _stream():
    # the raw cudaStream_t of torch's current stream on the current device (what torch.cuda.current_stream().cuda_stream
    # returns, without building a Stream object: ~10 us -> ~0.5 us, once per launch)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class KernelTimer:
    """Optional per-kernel-class CUDA-event timing on the launching (current) stream.  bench.py turns
    it on for the classes it reports a roofline for; off (the default) it costs one dict lookup."""

    def __init__(self):
        self.classes = set()
        self.events = {}

    def enable(self, classes):
        self.classes = set(classes)
        self.events = {c: [] for c in self.classes}

    def disable(self):
        self.classes = set()

    def start(self, cls):
        if cls not in self.classes:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (cls, e0, e1)

    def stop(self, tok, work=0.0):
        if tok is not None:
            tok[2].record()
            self.events[tok[0]].append((tok[1], tok[2], work))

    def summary(self):
        """{class: (launches, total_ms, total_work)} -- call after a synchronize."""
        return {c: (len(ev), sum(a.elapsed_time(b) for a, b, _ in ev), sum(w for _, _, w in ev))
                for c, ev in self.events.items()}


TIMER = KernelTimer()


def _call(name, *args):
    """one C-ABI call = one kernel launch; timed under its own name when bench.py asks for it"""
    tk = TIMER.start(name)
    L.check(getattr(L.lib(), name)(*args), name)
    TIMER.stop(tk)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.EmoError("emo ops need CUDA tensors: the hot path has no CPU implementation")


def embed_fwd(tok, seg, e_tok, e_seg, pe, out, scale, drop_p=0.0, seed=0, batch_first=True):
    """tok/seg int64 [B,T] (batch_first) or [T,B]; out [B,T,d]."""
    _need_cuda(tok, e_tok, out)
    if batch_first:
        B, T = tok.shape
        sb, st = tok.stride(0), tok.stride(1)
    else:
        T, B = tok.shape
        sb, st = tok.stride(1), tok.stride(0)
    if seg is not None:
        assert seg.stride() == tok.stride() and seg.dtype == torch.int64
    assert tok.dtype == torch.int64
    d = e_tok.shape[1]
    _call("emo_embed_fwd", _p(tok), _p(seg), sb, st, _p(e_tok), _p(e_seg), _p(pe), _p(out), B, T, d,
                                  float(scale), float(drop_p), int(seed), _dt(out), _stream())
    return out


def embed_rows(tok, seg, pos, e_tok, e_seg, pe, out, scale, advance_pos=None):
    """decode: tok/seg/pos int64 [rows] on the device; out [rows, d]; advance_pos (int64 [rows], may be `pos`)
    receives pos + 1."""
    _call("emo_embed_rows", _p(tok), _p(seg), _p(pos), _p(e_tok), _p(e_seg), _p(pe), _p(out), tok.shape[0],
          e_tok.shape[1], float(scale), _p(advance_pos), _dt(out), _stream())
    return out


def embed_bwd(tok, seg, dout, d_e_tok, d_e_seg, scale, drop_p=0.0, seed=0, pad_idx=-1, batch_first=True):
    if batch_first:
        B, T = tok.shape
        sb, st = tok.stride(0), tok.stride(1)
    else:
        T, B = tok.shape
        sb, st = tok.stride(1), tok.stride(0)
    d = d_e_tok.shape[1]
    _call("emo_embed_bwd", _p(tok), _p(seg), sb, st, _p(dout), _p(d_e_tok), _p(d_e_seg), B, T, d,
          float(scale), float(drop_p), int(seed), int(pad_idx), _dt(dout), _stream())


def ln_fwd(x, gamma, beta, y, mean=None, rstd=None, eps=1e-5):
    rows = x.numel() // x.shape[-1]
    _call("emo_ln_fwd", _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, x.shape[-1], eps,
                               _dt(x), _stream())
    return y


def ln_res_fwd(x, res, gamma, beta, y, mean=None, rstd=None, sum_out=None, eps=1e-5):
    """y = LN(x + res), the sum formed in fp32 in the kernel; sum_out (optional) = x + res for ln_bwd"""
    rows = x.numel() // x.shape[-1]
    _call("emo_ln_res_fwd", _p(x), _p(res), _p(sum_out), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows,
          x.shape[-1], eps, _dt(x), _stream())
    return y


def linear_ln_res_fwd(x, w, bias, res, gamma, beta, y, mean=None, rstd=None, sum_out=None, drop_p=0.0, seed=0, eps=1e-5):
    """y = LN(res + dropout(x . w^T + bias)) in one launch (emo_gemm_ln_res): bf16, w [512, K]; sum_out = the pre-LN sum"""
    _need_cuda(x, w, res, y)
    M, K = x.shape
    assert w.shape[0] == 512 and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    tk = TIMER.start("gemm_ln")          # its own class: gemm_ln_kernel is a projection AND a LayerNorm (HBM-bound half the time)
    L.check(L.lib().emo_gemm_ln_res(M, K, _p(x), x.stride(0), _p(w), w.stride(0), _p(bias), float(drop_p), int(seed), _p(res),
                                    res.stride(0), _p(gamma), _p(beta), float(eps), _p(y), y.stride(0), _p(sum_out),
                                    sum_out.stride(0) if sum_out is not None else 0, _p(mean), _p(rstd), _stream()), "emo_gemm_ln_res")
    TIMER.stop(tk, 2.0 * M * 512 * K)
    return y


def ln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, add_in=None, dx_drop=None, drop_p=0.0, seed=0, dxsum=None):
    """dxsum (fp32 [d]) += column sums of dx_drop (or dx): the bias gradient of the projection below."""
    rows = x.numel() // x.shape[-1]
    _call("emo_ln_bwd", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(add_in), _p(dx), _p(dx_drop),
                               float(drop_p), int(seed), _p(dgamma), _p(dbeta), _p(dxsum), rows, x.shape[-1], _dt(x),
                               _stream())


def dropout_apply(x, y, drop_p, seed):
    _call("emo_dropout_apply", _p(x), _p(y), x.numel(), float(drop_p), int(seed), _dt(x), _stream())
    return y


def gemm(op, M, N, K, A, lda, B, ldb, Cmat, ldc, bias=None, act=ACT_NONE, aux=None, aux_out=None, ld_aux=0,
         aux_scale=1.0, drop_p=0.0, seed=0, residual=None, ld_res=0, alpha=1.0, accumulate=False, colsum_out=None,
         ln=None):
    """Raw GEMM call; A/B/Cmat are tensors (or (tensor, element_offset) handled by the caller via views).
    colsum_out (fp32 [N]) += column sums of the stored C (bias gradient): fused into the epilogue on the
    bf16 tensor-core path, a separate emo_colsum launch otherwise.
    ln = (gamma, beta, out): A := LayerNorm(A) before the product (K == 512, bf16), normalised rows also written to
    `out` -- fused into the decode-rows kernel, a separate LN launch for larger M."""
    _need_cuda(A, B, Cmat)
    if ln is not None and (A.dtype != torch.bfloat16 or K != 512 or op == GEMM_TN):
        ln_fwd(A, ln[0], ln[1], ln[2])          # fp32 parity mode: LayerNorm as its own launch
        A, lda, ln = ln[2], ln[2].stride(0), None
    fuse_cs = (colsum_out is not None and A.dtype == torch.bfloat16 and Cmat.dtype == torch.bfloat16 and N % 64 == 0
               and ldc % 8 == 0 and Cmat.data_ptr() % 16 == 0 and not accumulate)
    e = L.Epilogue(_p(bias), act, _p(aux), _p(aux_out), ld_aux, aux_scale, drop_p, int(seed), _p(residual), ld_res,
                   alpha, 1 if accumulate else 0, None, _p(colsum_out) if fuse_cs else None,
                   _p(ln[0]) if ln else None, _p(ln[1]) if ln else None, _p(ln[2]) if ln else None,
                   ln[2].stride(0) if ln else 0)
    assert A.dtype == B.dtype
    tk = TIMER.start("gemm")
    L.check(L.lib().emo_gemm(op, M, N, K, _p(A), lda, _p(B), ldb, _p(Cmat), ldc, _dt(A), _dt(Cmat), C.byref(e),
                             _stream()), "emo_gemm")
    TIMER.stop(tk, 2.0 * M * N * K)
    if colsum_out is not None and not fuse_cs:
        colsum(Cmat, colsum_out, n=N)
    return Cmat


def linear_fwd(x, w, out, bias=None, **kw):
    """out[M,N] = x[M,K] . w[N,K]^T (+epilogue); 2-D row-major views (last stride 1)."""
    M, K = x.shape
    N = w.shape[0]
    return gemm(GEMM_NT, M, N, K, x, x.stride(0), w, w.stride(0), out, out.stride(0), bias=bias, **kw)


def linear_fwd_t(x, w_t, out, bias=None, **kw):
    """out[M,N] = x[M,K] . w_t[K,N]  (HF Conv1D weight layout [in,out])."""
    M, K = x.shape
    N = w_t.shape[1]
    return gemm(GEMM_NN, M, N, K, x, x.stride(0), w_t, w_t.stride(0), out, out.stride(0), bias=bias, **kw)


def linear_dgrad(dy, w, dx, **kw):
    """dx[M,K] = dy[M,N] . w[N,K]."""
    M, N = dy.shape
    K = w.shape[1]
    return gemm(GEMM_NN, M, K, N, dy, dy.stride(0), w, w.stride(0), dx, dx.stride(0), **kw)


def linear_dgrad_t(dy, w_t, dx, **kw):
    """dx[M,K] = dy[M,N] . w_t[K,N]^T  (Conv1D layout)."""
    M, N = dy.shape
    K = w_t.shape[0]
    return gemm(GEMM_NT, M, K, N, dy, dy.stride(0), w_t, w_t.stride(0), dx, dx.stride(0), **kw)


def linear_wgrad(dy, x, dw):
    """dw[N,K] (fp32) += dy[M,N]^T . x[M,K]."""
    M, N = dy.shape
    K = x.shape[1]
    return gemm(GEMM_TN, N, K, M, dy, dy.stride(0), x, x.stride(0), dw, dw.stride(0), accumulate=True)


def linear_wgrad_t(dy, x, dw_t):
    """dw_t[K,N] (fp32) += x[M,K]^T . dy[M,N]  (Conv1D layout)."""
    M, N = dy.shape
    K = x.shape[1]
    return gemm(GEMM_TN, K, N, M, x, x.stride(0), dy, dy.stride(0), dw_t, dw_t.stride(0), accumulate=True)


def colsum(x, out, n=None):
    M = x.shape[0]
    N = x.shape[1] if n is None else n
    _call("emo_colsum", _p(x), x.stride(0), M, N, _p(out), _dt(x), _stream())


def _favor_ld(q, k, v):
    """token stride of the [B,T,H,64] q/k/v views (size-1 dims carry arbitrary strides in torch)"""
    B, T, H, E = q.shape
    assert E == 64 and q.stride(3) == 1 and (H == 1 or q.stride(2) == 64)
    ld = q.stride(1) if T > 1 else (q.stride(0) if B > 1 else H * 64)
    assert B == 1 or T == 1 or q.stride(0) == T * ld, "q/k/v must be batch-contiguous views"
    for t in (k, v):
        assert t.shape == q.shape and t.stride(3) == 1
        assert (T == 1 or t.stride(1) == ld) and (B == 1 or t.stride(0) == q.stride(0))
    return ld


def _row_ld(t):
    """row stride of a [B,T,d] (or [rows,d]) view"""
    return t.stride(-2) if t.shape[-2] > 1 else t.shape[-1]


def favor_nseg(B, T, H, dtype):
    return int(L.lib().emo_favor_nseg(B, T, H, F32 if dtype == torch.float32 else BF16))


def favor_workspace(B, T, H, dtype, device):
    """[B,H,nseg+1,128,80] fp32 segment-state workspace for favor_fwd(seg_states=...) / favor_bwd
    (emo_favor_nseg returns the slot count: nseg local sums, turned into prefixes in place, + the total)"""
    return torch.empty(B, H, favor_nseg(B, T, H, dtype), 128, 80, dtype=torch.float32, device=device)


def favor_fwd(q, k, v, omega, out, den=None, state_out=None, state_in=None, seg_states=None):
    """q,k,v: [B,T,H,64] views with a common token stride; out [B,T,H*64] view.  seg_states (from
    favor_workspace) switches on the segment-parallel schedule and is what favor_bwd consumes."""
    B, T, H, E = q.shape
    ld = _favor_ld(q, k, v)
    assert seg_states is None or state_in is None
    tk = TIMER.start("favor_fwd")
    L.check(L.lib().emo_favor_fwd(_p(q), _p(k), _p(v), ld, _p(omega), _p(out), _row_ld(out), _p(den),
                                  _p(state_in), _p(state_out), _p(seg_states), B, T, H, _dt(q), _stream()), "emo_favor_fwd")
    TIMER.stop(tk, float(B * T * H * 64 * 4 * q.element_size()))      # algorithmic bytes: read q,k,v + write out
    return out


def favor_bwd(q, k, v, omega, out, dout, den, seg_states, dq, dk, dv, seg_rstates=None):
    B, T, H, E = q.shape
    ld = _favor_ld(q, k, v)
    ldd = _favor_ld(dq, dk, dv)
    if seg_rstates is None:
        seg_rstates = torch.empty_like(seg_states)
    tk = TIMER.start("favor_bwd")
    L.check(L.lib().emo_favor_bwd(_p(q), _p(k), _p(v), ld, _p(omega), _p(out), _p(dout), _row_ld(out),
                                  _p(den), _p(seg_states), _p(seg_rstates), _p(dq), _p(dk), _p(dv), ldd, B, T, H, _dt(q),
                                  _stream()), "emo_favor_bwd")
    TIMER.stop(tk, float(B * T * H * 64 * 8 * q.element_size()))      # read q,k,v,out,dout + write dq,dk,dv


def favor_step(q, k, v, omega, state, out):
    """q,k,v: [B,H,64] views (sequence stride = stride(0)); state [B,H,128,80] fp32; out [B,H*64]."""
    B, H, E = q.shape
    _call("emo_favor_step", _p(q), _p(k), _p(v), q.stride(0), _p(omega), _p(state), _p(out), out.stride(0),
                                   B, H, _dt(q), _stream())
    return out


def _qkv_ld(t):
    """token stride of a [B,T,H,64] view, and check of the batch stride"""
    B, T, H, E = t.shape
    assert E == 64 and t.stride(3) == 1 and (H == 1 or t.stride(2) == 64)
    ld = t.stride(1) if T > 1 else (t.stride(0) if B > 1 else H * 64)
    assert B == 1 or T == 1 or t.stride(0) == T * ld, "attention operands must be batch-contiguous views"
    return ld


def attn_fwd(q, k, v, out, lse, scale, drop_p=0.0, seed=0):
    """q [B,Tq,H,64], k,v [B,Tk,H,64] views; out [B,Tq,H*64]; lse [B,H,Tq] fp32."""
    B, Tq, H, E = q.shape
    Tk = k.shape[1]
    _call("emo_attn_fwd", _p(q), _p(k), _p(v), _qkv_ld(q), _qkv_ld(k), _p(out), _row_ld(out), _p(lse),
          B, Tq, Tk, H, scale, drop_p, int(seed), _dt(q), _stream())
    return out


def attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, scale, drop_p=0.0, seed=0):
    B, Tq, H, E = q.shape
    Tk = k.shape[1]
    _call("emo_attn_bwd", _p(q), _p(k), _p(v), _qkv_ld(q), _qkv_ld(k), _p(out), _p(dout), _row_ld(out),
          _p(lse), _p(dq), _p(dk), _p(dv), _qkv_ld(dq), _qkv_ld(dk), B, Tq, Tk, H, scale,
          drop_p, int(seed), _dt(q), _stream())


def attn_decode_step(qkv, kv_cache, pos, out, scale):
    """qkv [B, 3*H*64] (the new token's q | k | v), kv_cache [B, max_len, 2*H*64], pos int64 [B] (device): appends k | v
    at pos[b], attends over keys 0..pos[b]; out [B, H*64]"""
    B, max_len = kv_cache.shape[0], kv_cache.shape[1]
    H = kv_cache.shape[2] // 128
    _call("emo_attn_decode_step", _p(qkv), qkv.stride(0), _p(kv_cache), max_len, _p(pos), _p(out), out.stride(0), B, H,
          float(scale), _dt(qkv), _stream())
    return out


def relattn_decode_step(qkv, kv_cache, pos, rtab, r_w_bias, r_r_bias, mem_len, out, scale):
    """stage-1 decode: qkv [B, 3*H*64] of the new token, kv_cache [B, cap, 2*H*64], pos int64 [B] (device), rtab
    [mem_len + 1, H*64] indexed by distance; attends over the last mem_len + 1 positions; out [B, H*64]"""
    B, cap = kv_cache.shape[0], kv_cache.shape[1]
    H = kv_cache.shape[2] // 128
    _call("emo_relattn_decode_step", _p(qkv), qkv.stride(0), _p(kv_cache), cap, _p(pos), _p(rtab), _p(r_w_bias), _p(r_r_bias),
          int(mem_len), _p(out), out.stride(0), B, H, float(scale), _dt(qkv), _stream())
    return out


def relattn_fwd(q, k, v, r, r_w_bias, r_r_bias, out, lse, scale, drop_p=0.0, seed=0):
    """r [Tk,H,64] (row p = distance Tk-1-p); biases [H,64] fp32."""
    B, Tq, H, E = q.shape
    Tk = k.shape[1]
    _call("emo_relattn_fwd", _p(q), _p(k), _p(v), _qkv_ld(q), _qkv_ld(k), _p(r), r.stride(0) if Tk > 1 else H * 64,
          _p(r_w_bias), _p(r_r_bias), _p(out), _row_ld(out), _p(lse), B, Tq, Tk, H, scale, drop_p, int(seed),
          _dt(q), _stream())
    return out


def relattn_bwd(q, k, v, r, r_w_bias, r_r_bias, out, dout, lse, dq, dk, dv, dr, d_rw, d_rr, scale, drop_p=0.0, seed=0):
    B, Tq, H, E = q.shape
    Tk = k.shape[1]
    _call("emo_relattn_bwd", _p(q), _p(k), _p(v), _qkv_ld(q), _qkv_ld(k), _p(r), r.stride(0) if Tk > 1 else H * 64,
          _p(r_w_bias), _p(r_r_bias), _p(out), _p(dout), _row_ld(out), _p(lse), _p(dq), _p(dk), _p(dv),
          _qkv_ld(dq), _qkv_ld(dk), _p(dr), _p(d_rw), _p(d_rr), B, Tq, Tk, H, scale, drop_p, int(seed),
          _dt(q), _stream())


def _tgt_layout(tgt, batch_first):
    """rows are enumerated batch-major (r = b*T + t) to match the [B,T,.] activations."""
    if batch_first:
        B, T = tgt.shape
        return B * T, T, tgt.stride(0), tgt.stride(1)
    T, B = tgt.shape
    return B * T, T, tgt.stride(1), tgt.stride(0)


def ce_count(tgt, ignore_index, count, batch_first=True):
    rows, inner, so, si = _tgt_layout(tgt, batch_first)
    _call("emo_ce_count", _p(tgt), rows, inner, so, si, int(ignore_index), _p(count), _stream())


def ce_fwd_bwd(logits, tgt, V, ignore_index, count, loss_sum, ncorrect=None, pred=None, dlogits=None, gscale=1.0,
               batch_first=True):
    """logits fp32 [rows, ld] (rows batch-major); dlogits [rows, ld_dl] or None."""
    rows, inner, so, si = _tgt_layout(tgt, batch_first)
    assert logits.dtype == torch.float32 and logits.shape[0] == rows
    _call("emo_ce_fwd_bwd", _p(logits), logits.stride(0), _p(tgt), rows, inner, so, si, V, int(ignore_index),
                                   _p(count), float(gscale), _p(loss_sum), _p(ncorrect), _p(pred), _p(dlogits),
                                   0 if dlogits is None else dlogits.stride(0),
                                   F32 if dlogits is None else _dt(dlogits), _stream())


def sumsq(g, out):
    _call("emo_sumsq", _p(g), g.numel(), _p(out), _stream())


def adam_step(p, g, m, v, p_bf16, lr, beta1, beta2, eps, step, gnorm_sq=None, max_norm=0.0, grad_scale=1.0,
              zero_grad=False):
    _call("emo_adam_step", _p(p), _p(g), _p(m), _p(v), _p(p_bf16), p.numel(), lr, beta1, beta2, eps,
                                  int(step), _p(gnorm_sq), float(max_norm), float(grad_scale), 1 if zero_grad else 0,
                                  _stream())


def cast(src, dst):
    _call("emo_cast", _p(src), _p(dst), src.numel(), _dt(src), _dt(dst), _stream())
    return dst


def sample(logits, V, temperature, top_p, u, out, status=None, greedy=False, banned=None):
    """banned: uint8 [rows, V] (non-zero = inadmissible token) or None"""
    rows = logits.shape[0]
    if banned is not None:
        assert banned.dtype == torch.uint8 and banned.shape[-1] == V and banned.is_contiguous()
    if torch.is_tensor(temperature):                    # fp32 [rows] on the device: one temperature per row
        assert temperature.dtype == torch.float32 and temperature.numel() >= rows and temperature.is_cuda
        _call("emo_sample_rows", _p(logits), logits.stride(0), rows, V, _p(temperature), float(top_p), _p(u),
              1 if greedy else 0, _p(out), _p(status), _p(banned), _stream())
        return out
    _call("emo_sample", _p(logits), logits.stride(0), rows, V, float(temperature), float(top_p), _p(u),
                               1 if greedy else 0, _p(out), _p(status), _p(banned), _stream())
    return out


def logits_sample(x, w, bias, logits, V, temperature, top_p, u, out, status=None, greedy=False, banned=None, ln=None):
    """Fused logits projection + sampler (emo_logits_sample): logits[:, :V] = LN?(x) . w[:V]^T + bias, then the draw of
    `sample` from an on-chip copy.  x bf16 [rows, 512]; w bf16 [V, 512]; logits fp32 [rows, ld]; ln = (gamma, beta)."""
    _need_cuda(x, w, logits)
    rows = x.shape[0]
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and logits.dtype == torch.float32
    if banned is not None:
        assert banned.dtype == torch.uint8 and banned.shape[-1] == V and banned.is_contiguous()
    t_rows = None
    if torch.is_tensor(temperature):
        assert temperature.dtype == torch.float32 and temperature.numel() >= rows and temperature.is_cuda
        t_rows, temperature = temperature, 1.0
    _call("emo_logits_sample", _p(x), x.stride(0), _p(ln[0]) if ln else None, _p(ln[1]) if ln else None, _p(w), w.stride(0),
          _p(bias), rows, V, x.shape[1], _p(logits), logits.stride(0), float(temperature), _p(t_rows), float(top_p), _p(u),
          1 if greedy else 0, _p(out), _p(status), _p(banned), _stream())
    return out


def performer_decode_step(w_bf16, w_f32, layer_offs, off_tok, off_seg, off_outw, off_outb, pe, omegas, state, tok, seg, pos,
                          scratch, logits, n_layer, batch, n_token, emb_scale):
    """one cooperative kernel = one decode step of the stage-2 Performer (include/emo_b200.h)"""
    _call("emo_performer_decode_step", _p(w_bf16), _p(w_f32), _p(layer_offs), int(off_tok), int(off_seg), int(off_outw),
          int(off_outb), _p(pe), _p(omegas), _p(state), _p(tok), _p(seg), _p(pos), _p(scratch), _p(logits), n_layer, batch,
          n_token, logits.stride(0), float(emb_scale), _stream())
