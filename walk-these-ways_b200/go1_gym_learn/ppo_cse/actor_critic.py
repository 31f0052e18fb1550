"""ActorCritic of ppo_cse (reference go1_gym_learn/ppo_cse/actor_critic.py:19-147) on hand-written kernels.

Same public surface (AC_Args, ActorCritic(num_obs, num_privileged_obs, num_obs_history, num_actions),
.adaptation_module / .actor_body / .critic_body as nn.Sequential so checkpoints and TorchScript exports keep
the reference's names, .act / .evaluate / .act_student / .act_teacher / .get_actions_log_prob / .action_mean /
.action_std / .entropy), but:
  * all parameters are views into ONE flat fp32 buffer (adaptation module first), gradients into one flat
    gradient buffer -> one grad-norm, one Adam launch, one NCCL all-reduce per optimizer step;
  * forward and backward are explicit go1_gemm calls (fp32 CUDA-core or tcgen05 TF32) with fused
    bias+ELU epilogues; cat(obs_history, latent) is never materialised (the 2 extra input columns are a
    second, K=2 GEMM accumulated into the first layer's pre-activation);
  * no autograd graph: the backward pass is written out (see `backward_ppo`, `backward_adaptation`).
"""
import torch
import torch.nn as nn
from params_proto import PrefixProto

from go1_b200 import capi


class AC_Args(PrefixProto, cli=False):
    # policy
    init_noise_std = 1.0
    actor_hidden_dims = [512, 256, 128]
    critic_hidden_dims = [512, 256, 128]
    activation = 'elu'  # only elu runs on the fused kernels
    adaptation_module_branch_hidden_dims = [256, 128]
    use_decoder = False
    gemm_impl = 1       # 1 = tcgen05 TF32 tensor cores (default; torch 1.10, the reference's pin, also ran these matmuls in TF32), 0 = fp32 CUDA cores (exact)


def _mlp(in_dim, hidden, out_dim):
    layers, d = [], in_dim
    for h in hidden:
        layers += [nn.Linear(d, h), nn.ELU()]
        d = h
    layers.append(nn.Linear(d, out_dim))
    return nn.Sequential(*layers)


def _empty(*shape, device):
    """Scratch buffers outlive the Runner's torch.inference_mode() rollout block and are written again by the update,
    so they must be ordinary (non-inference) tensors whichever mode they are first needed in."""
    with torch.inference_mode(False):
        return torch.empty(*shape, device=device)


def _aligned(t_or_ptr, ld):
    p = t_or_ptr if isinstance(t_or_ptr, int) else t_or_ptr.data_ptr()
    return (p & 15) == 0 and (ld & 3) == 0


class _Net:
    """Forward/backward of one MLP whose first layer reads [x (K0 columns) | extra (E columns)].

    impl 0: every product is one fp32 CUDA-core go1_gemm.  impl 1: the large products run on the tcgen05 TF32 kernel,
    which reads its operands through TMA (16-byte aligned rows) in either major: forward K-major, dgrad with W as an
    MN-major B operand, wgrad with dz and the layer input as MN-major A and B operands -- no transposed copies.  Only
    the first-layer weight block W[:, :K0] (row stride 2102 floats) is packed to a TMA-readable copy, cached per
    weight version."""

    def __init__(self, seq, flat, grad, offsets, owner):
        self.linears = [m for m in seq if isinstance(m, nn.Linear)]
        self.specs = []                       # (w_off, b_off, out, in)
        for lin in self.linears:
            o, i = lin.weight.shape
            self.specs.append((offsets[id(lin.weight)], offsets[id(lin.bias)], o, i))
        last = self.linears[-1]
        self.end = offsets[id(last.bias)] + last.bias.numel()
        self.flat, self.grad, self.owner = flat, grad, owner
        self.acts = {}
        self._cache = {}
        self._ep = capi.Go1GemmEpilogue()

    def _buf(self, key, M, width):
        t = self.acts.get(key)
        if t is None or t.shape[0] < M or t.shape[1] != width:
            t = _empty(M, width, device=self.flat.device)
            self.acts[key] = t
        return t[:M]

    def _cached(self, key, build):
        """A packed copy of weights, rebuilt when the weights changed (weights_version).  The entry keeps its builder so that
        ActorCritic.ensure_packed() can refresh every copy eagerly before a CUDA graph that reads them is replayed: the graphs contain
        no packing kernels (the rollout replays one 24 times per weight version)."""
        ver = self.owner.weights_version
        hit = self._cache.get(key)
        if hit is None or hit[0] != ver:
            if torch.cuda.is_current_stream_capturing():
                raise capi.Go1Error("stale packed weights during graph capture: call ActorCritic.ensure_packed() first")
            hit = (ver, build(hit[1] if hit else None), build)
            self._cache[key] = hit
        return hit[1]

    def refresh_packed(self):
        ver = self.owner.weights_version
        for key, hit in list(self._cache.items()):
            if hit[0] != ver:
                self._cache[key] = (ver, hit[2](hit[1]), hit[2])

    @staticmethod
    def _p(x):
        return x.data_ptr() if torch.is_tensor(x) else x

    def _gemm(self, ta, tb, M, N, K, A, lda, B, ldb, Cm, ldc, bias=None, act=0, acc=0, impl=0, extra=None, w_extra=0, ld_w_extra=0, dact_y=None, lead_cols=0,
              colsum=None, bwd_extra=None):
        ep = self._ep
        ep.lead_cols = lead_cols
        ep.colsum = self._p(colsum) if colsum is not None else None
        if bwd_extra is not None:      # (extra [M][E], w_extra ptr, ld, g_w_extra ptr, ld, dextra [M][E] or None)
            ex, wex, ldw, gw, ldg, dex = bwd_extra
            ep.bwd_extra, ep.ld_bwd_extra, ep.num_bwd_extra = ex.data_ptr(), ex.stride(0), ex.shape[1]
            ep.bwd_w_extra, ep.ld_bwd_w_extra, ep.g_w_extra, ep.ld_g_w_extra = wex, ldw, gw, ldg      # gw None: no weight-gradient reduction
            ep.d_extra, ep.ld_d_extra = (dex.data_ptr(), dex.stride(0)) if dex is not None else (None, 0)
        else:
            ep.num_bwd_extra = 0
        ep.bias = self._p(bias) if bias is not None else None
        ep.act, ep.accumulate = act, acc
        if extra is not None:
            ep.extra, ep.ld_extra, ep.w_extra, ep.ld_w_extra, ep.num_extra = extra.data_ptr(), extra.stride(0), w_extra, ld_w_extra, extra.shape[1]
        else:
            ep.extra, ep.num_extra = None, 0
        if dact_y is not None:
            ep.dact_y, ep.ld_dact_y = dact_y.data_ptr(), dact_y.stride(0)
        else:
            ep.dact_y = None
        capi.check(capi.lib().go1_gemm_ex(ta, tb, M, N, K, self._p(A), lda, self._p(B), ldb, self._p(Cm), ldc, ep, impl, capi.stream_ptr()), "go1_gemm")

    @staticmethod
    def _tma_ok(x, ld):
        """TMA-readable K-major operand: 16-byte aligned base, row stride a multiple of 16 bytes."""
        p = x.data_ptr() if torch.is_tensor(x) else x
        return (p & 15) == 0 and (ld & 3) == 0

    def forward(self, x, ldx, K0, extra, M, impl, tag="a", first_out=None):
        """x: [M][K0] rows with stride ldx; extra: [M][E] contiguous or None. Returns list of layer outputs.
        first_out: the first layer's activated output if the caller already produced it (ActorCritic.forward_all)."""
        outs, inp, ld_in = [], x, ldx
        n = len(self.specs)
        for li, (wo, bo, o, i) in enumerate(self.specs):
            if li == 0 and first_out is not None:
                outs.append(first_out)
                inp, ld_in = first_out, first_out.stride(0)
                continue
            if li == 1 and impl == 1 and self._tail_ok(inp, ld_in, M):
                return outs + self._forward_tail(inp, ld_in, M, tag)
            y = self._buf((tag, li), M, o)
            W = self.flat[wo:wo + o * i]
            b = self.flat[bo:bo + o]
            act = 1 if li < n - 1 else 0
            first_extra = li == 0 and extra is not None
            if impl == 1 and li == n - 1 and li > 0 and o <= 16 and i % 4 == 0 and i <= 512 and self._tma_ok(inp, ld_in):
                # the narrow head: one bandwidth-bound pass instead of a padded tensor-core tile
                capi.check(capi.lib().go1_skinny_forward(self._p(inp), ld_in, W.data_ptr(), i, b.data_ptr(), capi.ptr(y), o, M, o, i, capi.stream_ptr()), "skinny_forward")
                outs.append(y)
                continue
            K = K0 if first_extra else i
            Wm, ldw = W, i
            tc = impl == 1 and self._tma_ok(inp, ld_in)
            if tc and (not self._tma_ok(W, i) or (li == 0 and K >= 1024 and i % 32 != 0)):
                # pack W[:, :K] into a TMA-readable copy whose row pitch is a multiple of 128 bytes (K % 4 == 0); for the long first-layer
                # rows the aligned pitch alone is worth 30 % (misaligned 128-byte box rows cost an extra L2 sector each)
                if K % 4 == 0:
                    KPk = (K + 31) // 32 * 32
                    Wm = self._cached(("pack", li), lambda old, W=W, o=o, i=i, K=K, KPk=KPk: (old if old is not None else _empty(o, KPk, device=W.device)[:, :K]).copy_(W.view(o, i)[:, :K]))
                    ldw = KPk
                else:
                    tc = False
            if first_extra:     # y = act(x W[:, :K0]^T + extra W[:, K0:]^T + b): the 2 trailing columns ride in the epilogue
                self._gemm(0, 1, M, o, K0, inp, ld_in, Wm, ldw, y, o, b, act, 0, 1 if tc else 0, extra=extra, w_extra=W.data_ptr() + 4 * K0, ld_w_extra=i)
            else:
                self._gemm(0, 1, M, o, K, inp, ld_in, Wm, ldw, y, o, b, act, 0, 1 if tc else 0)
            outs.append(y)
            inp, ld_in = y, y.stride(0)
        return outs

    _TAILS = {(512, 256, 128): 3, (256, 128): 2}       # hidden widths behind the first layer that go1_mlp_tail_forward fuses

    def _tail_ok(self, x1, ldx1, M):
        """The layers behind the first one form a tail the fused kernel supports (and its TMA operands are aligned)."""
        if not self.owner.fuse_tail:
            return False
        widths = tuple(sp[2] for sp in self.specs[:-1])
        if self._TAILS.get(widths) != len(self.specs) - 1 or self.specs[-1][2] > 12:
            return False
        ok = self._tma_ok(x1, ldx1)
        for wo, bo, o, i in self.specs[1:-1]:
            ok = ok and ((self.flat.data_ptr() + 4 * wo) & 15) == 0 and i % 4 == 0
        return ok

    def _tail_problem(self, x1, ldx1, M, tag):
        """(Go1TailProblem, [outputs of layers 1..]) of this net's tail for go1_mlp_tail_forward_grouped."""
        sp, flat = self.specs, self.flat
        ptr = lambda off: flat.data_ptr() + 4 * off
        q = capi.Go1TailProblem()
        (w2, b2, n2, k1) = sp[1]
        y2 = self._buf((tag, 1), M, n2)
        q.x, q.ldx, q.W2, q.b2, q.y2, q.ldy2 = self._p(x1), ldx1, ptr(w2), ptr(b2), y2.data_ptr(), y2.stride(0)
        if len(sp) == 4:
            (w3, b3, n3, _), (wh, bh, nh, _) = sp[2], sp[3]
            y3 = self._buf((tag, 2), M, n3)
            out = self._buf((tag, 3), M, nh)
            q.W3, q.b3, q.y3, q.ldy3 = ptr(w3), ptr(b3), y3.data_ptr(), y3.stride(0)
            outs = [y2, y3, out]
        else:
            (wh, bh, nh, _) = sp[2]
            n3 = 0
            out = self._buf((tag, 2), M, nh)
            q.W3, q.b3, q.y3, q.ldy3 = None, None, None, 0
            outs = [y2, out]
        q.Wh, q.bh, q.nh, q.out, q.ldout = ptr(wh), ptr(bh), nh, out.data_ptr(), out.stride(0)
        return q, outs, (k1, n2, n3)

    def _forward_tail(self, x1, ldx1, M, tag):
        """Layers 1.. (and the head) in ONE launch; returns their outputs in layer order."""
        q, outs, (k1, n2, n3) = self._tail_problem(x1, ldx1, M, tag)
        arr = (capi.Go1TailProblem * 1)(q)
        capi.check(capi.lib().go1_mlp_tail_forward_grouped(arr, 1, M, k1, n2, n3, capi.stream_ptr()), "go1_mlp_tail_forward")
        return outs

    def tail_bwd_ok(self, outs, dout):
        """This body is a ...-256-128-head one whose backward first half go1_mlp_tail_backward_grouped supports (launches nothing)."""
        sp = self.specs
        if len(sp) != 4 or sp[2][2] != 128 or sp[2][3] != 256 or sp[1][2] != 256 or sp[3][2] > 12 or not self.owner.grads_prezeroed:
            return False
        y3, y2 = outs[2], outs[1]
        return self._tma_ok(y3, y3.stride(0)) and self._tma_ok(y2, y2.stride(0)) and ((self.flat.data_ptr() + 4 * sp[2][0]) & 15) == 0 and dout.stride(1) == 1

    def tail_bwd_problem(self, outs, dout, M, tag):
        """Launches the head's wgrad (+ bias gradient) of this body and returns (Go1TailBwdProblem, dz3, dz2) for the fused backward first
        half (go1_mlp_tail_backward_grouped).  Call only if tail_bwd_ok()."""
        sp = self.specs
        (wo3, bo3, n3, n2), (woh, boh, nh, _), (wo2, bo2, _, _) = sp[2], sp[3], sp[1]
        y3, y2 = outs[2], outs[1]
        L, st = capi.lib(), capi.stream_ptr()
        gWh, gbh = self.grad[woh:woh + nh * n3], self.grad[boh:boh + nh]
        capi.check(L.go1_skinny_wgrad_ex(capi.ptr(dout), dout.stride(0), capi.ptr(y3), y3.stride(0), gWh.data_ptr(), n3, gbh.data_ptr(), M, nh, n3, 1, st), "skinny_wgrad")
        dz3, dz2 = self._buf((tag, "d", 2), M, n3), self._buf((tag, "d", 1), M, n2)
        q = capi.Go1TailBwdProblem()
        q.dout, q.lddout, q.nh, q.Wh = dout.data_ptr(), dout.stride(0), nh, self.flat.data_ptr() + 4 * woh
        q.y3, q.ldy3, q.W3, q.y2, q.ldy2 = y3.data_ptr(), y3.stride(0), self.flat.data_ptr() + 4 * wo3, y2.data_ptr(), y2.stride(0)
        q.dz3, q.lddz3, q.dz2, q.lddz2 = dz3.data_ptr(), dz3.stride(0), dz2.data_ptr(), dz2.stride(0)
        q.gb3, q.gb2 = self.grad.data_ptr() + 4 * bo3, self.grad.data_ptr() + 4 * bo2
        return q, dz3, dz2

    def backward(self, x, ldx, K0, extra, outs, dout, M, impl, accumulate, want_dextra=False, tag="a", dz1_out=None, aug_first=False, pre=None):
        """dout: gradient w.r.t. the network output [M][out] (the last layer has no activation).  Writes weight/bias grads
        into the flat grad buffer.  dz of every hidden layer comes out of the dgrad GEMM already multiplied by ELU'
        (fused epilogue).  dz1_out: optional [M][o1] strided view; when given the first layer's dz is written there and its wgrad is
        left to the caller (ActorCritic fuses the three first-layer wgrads into one GEMM).  aug_first: the caller's fused wgrad also yields the first
        layer's bias gradient and trailing-input weight gradients (augmented input columns), so the dgrad epilogue that produces the first layer's
        dz reduces neither of them (only d(extra) if requested).  Returns d(extra) [M][E] if requested."""
        L, st = capi.lib(), capi.stream_ptr()
        n = len(self.specs)
        dz = dout
        dextra = None
        bias_done = False      # this layer's bias gradient was already reduced in the epilogue of the dgrad product that made its dz
        extra_done = False     # likewise the trailing-input gradients of the first layer
        start = n - 1
        if pre is not None:    # (dz of layer n-2, dz of layer n-3) from go1_mlp_tail_backward_grouped, which also reduced their bias gradients;
            dz, start, bias_done = pre[0], n - 2, True      # the head's wgrad was launched by tail_bwd_problem
        for li in range(start, -1, -1):
            wo, bo, o, i = self.specs[li]
            W = self.flat[wo:wo + o * i]
            gW, gb = self.grad[wo:wo + o * i], self.grad[bo:bo + o]
            ldz = dz.stride(0)
            if li == 0:
                inp, ld_in, K = x, ldx, (K0 if extra is not None else i)
            else:
                inp, ld_in, K = outs[li - 1], outs[li - 1].stride(0), i
            prez = 1 if (accumulate or self.owner.grads_prezeroed) else 0      # the caller zeroed the gradient buffer: accumulate, no memsets
            skinny_w = not (li == 0 and dz1_out is not None) and o <= 16 and K >= 32
            if not bias_done and skinny_w and K % 4 == 0 and self._tma_ok(inp, ld_in):
                # the narrow heads: weight AND bias gradient in one bandwidth-bound pass over the layer input
                capi.check(L.go1_skinny_wgrad_ex(capi.ptr(dz), ldz, capi.ptr(inp), ld_in, gW.data_ptr(), i, gb.data_ptr(), M, o, K, prez, st), "skinny_wgrad")
                skinny_w, bias_done = False, True
            if not bias_done:
                capi.check(L.go1_colsum(capi.ptr(dz), ldz, capi.ptr(gb), M, o, accumulate, st), "colsum")
            bias_done = False
            # ---- wgrad: dW[o][K] = dz^T[o][M] inp[M][K]
            if li == 0 and dz1_out is not None:
                pass                                    # fused by the caller
            elif o <= 16 and K >= 32:   # the narrow heads: one bandwidth-bound pass instead of a padded GEMM tile
                if skinny_w:
                    capi.check(L.go1_skinny_wgrad(capi.ptr(dz), ldz, capi.ptr(inp), ld_in, gW.data_ptr(), i, M, o, K, prez, st), "skinny_wgrad")
            else:       # impl 1: both operands MN-major, read in place by the tcgen05 kernel
                tc = impl == 1 and M >= 64 and K >= 8 and self._tma_ok(dz, ldz) and self._tma_ok(inp, ld_in)
                # into a gradient buffer the caller has already zeroed the split-K partial tiles can accumulate directly (no zeroing pass)
                acc_w = 1 if (accumulate or (tc and self.owner.grads_prezeroed)) else 0
                q = self.owner._wgrad_queue
                if q is not None and tc and acc_w:      # deferred: ActorCritic launches the equal-shape wgrads of its MLPs as grouped products
                    q.append((o, K, M, ldz, ld_in, i, dz, inp, gW))
                else:
                    self._gemm(1, 0, o, K, M, dz, ldz, inp, ld_in, gW, i, None, 0, acc_w, 1 if tc else 0)
            if li == 0 and extra is not None and not extra_done:
                E = i - K0
                if want_dextra:
                    dextra = self._buf((tag, "dextra"), M, E)
                capi.check(L.go1_mlp_extra_backward(capi.ptr(dz), ldz, capi.ptr(extra), extra.stride(0), W.data_ptr() + 4 * K0, i, gW.data_ptr() + 4 * K0, i,
                                                    capi.ptr(dextra) if want_dextra else None, E, M, o, E, accumulate, st), "extra_backward")
            # ---- dgrad (+ fused ELU'): dz_prev[M][i] = (dz[M][o] W[o][i]) * ELU'(y_prev)
            if pre is not None and li == n - 2:
                dz, bias_done = pre[1], True            # produced (with its bias gradient) by the fused kernel
                continue
            if li > 0:
                dprev = dz1_out if (li == 1 and dz1_out is not None) else self._buf((tag, "d", li - 1), M, i)
                ldp = dprev.stride(0)
                yprev = outs[li - 1]
                if impl == 1 and self._tma_ok(dz, ldz) and self._tma_ok(W, i) and M >= 64:      # (the 12-wide actor head included: 19 us here, 22 us on the skinny pass)
                    # W read MN-major in place; the bias gradient of layer li-1 (column sums of dprev) rides in the epilogue
                    pwo, pbo, po, pi = self.specs[li - 1]
                    gb_prev = self.grad[pbo:pbo + po]
                    fuse = self.owner.fuse_bias_grad
                    if fuse and not accumulate and not self.owner.grads_prezeroed:
                        gb_prev.zero_()
                    bx = None
                    aug = aug_first and li == 1
                    if aug:
                        # the first layer's bias and trailing-input weight gradients come out of the caller's fused wgrad; only d(extra) is left
                        if extra is not None and want_dextra:
                            E0 = pi - K0
                            Wp = self.flat[pwo:pwo + po * pi]
                            dextra = self._buf((tag, "dextra"), M, E0)
                            dextra.zero_()
                            bx = (extra, Wp.data_ptr() + 4 * K0, pi, None, 0, dextra)
                        extra_done = True
                    elif fuse and li == 1 and extra is not None and self.owner.grads_prezeroed and not accumulate and 1 <= pi - K0 <= 4:
                        # dprev is the first layer's dz: its trailing-input weight gradient (and d(extra)) are reduced in this epilogue too
                        E0 = pi - K0
                        Wp = self.flat[pwo:pwo + po * pi]
                        gWp = self.grad[pwo:pwo + po * pi]
                        if want_dextra:
                            dextra = self._buf((tag, "dextra"), M, E0)
                            dextra.zero_()
                        bx = (extra, Wp.data_ptr() + 4 * K0, pi, gWp.data_ptr() + 4 * K0, pi, dextra if want_dextra else None)
                        extra_done = True
                    self._gemm(0, 0, M, i, o, dz, ldz, W, i, dprev, ldp, None, 2, 0, 1, dact_y=yprev, colsum=gb_prev if (fuse and not aug) else None, bwd_extra=bx)
                    bias_done = bool(fuse) or aug
                elif aug_first and li == 1:
                    raise capi.Go1Error("augmented first-layer wgrad: the dgrad that produces the first layer's dz must be a tcgen05 product")
                elif o <= 16:
                    pwo, pbo, po, pi = self.specs[li - 1]
                    gb_prev = self.grad[pbo:pbo + po]
                    fuse = impl == 1 and self.owner.fuse_bias_grad and i % 4 == 0 and self._tma_ok(W, i) and self._tma_ok(dprev, ldp) and self._tma_ok(yprev, yprev.stride(0))
                    if fuse and not accumulate and not self.owner.grads_prezeroed:
                        gb_prev.zero_()
                    # the bias gradient of layer li-1 (column sums of dprev) is reduced in the same pass
                    capi.check(L.go1_skinny_dgrad_ex(capi.ptr(dz), ldz, capi.ptr(W), i, capi.ptr(yprev), yprev.stride(0), capi.ptr(dprev), ldp,
                                                     gb_prev.data_ptr() if fuse else None, M, o, i, st), "skinny_dgrad")
                    bias_done = bool(fuse)
                else:
                    self._gemm(0, 0, M, i, o, dz, ldz, W, i, dprev, ldp, None, 2, 0, 0, dact_y=yprev)
                dz = dprev
        return dextra


class ActorCritic(nn.Module):
    is_recurrent = False
    HEAD = 16          # floats reserved in front of the flat parameter / gradient buffers (see flatten())

    def __init__(self, num_obs, num_privileged_obs, num_obs_history, num_actions, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str([key for key in kwargs.keys()]))
        self.decoder = AC_Args.use_decoder
        super().__init__()
        if AC_Args.activation != 'elu':
            raise NotImplementedError("the fused MLP kernels implement ELU (the reference's configured activation)")
        self.num_obs_history, self.num_privileged_obs, self.num_actions = num_obs_history, num_privileged_obs, num_actions
        self.adaptation_module = _mlp(num_obs_history, AC_Args.adaptation_module_branch_hidden_dims, num_privileged_obs)
        self.actor_body = _mlp(num_privileged_obs + num_obs_history, AC_Args.actor_hidden_dims, num_actions)
        self.critic_body = _mlp(num_privileged_obs + num_obs_history, AC_Args.critic_hidden_dims, 1)
        self.std = nn.Parameter(AC_Args.init_noise_std * torch.ones(num_actions))
        self.distribution = None
        self._flat = self._grad = None
        self._mean = self._value = self._logp = None
        self._sample_counter = 0
        self._counter_dev = None
        self.force_repack = False     # (kept for callers that set it; packing is never part of a graph any more, see ensure_packed)
        self._packed_version = -1
        self._wgrad_queue = None      # list while backward_ppo collects the tensor-core wgrads of the layers behind the first ones
        self.sample_seed = 0
        self.injected_eps = None      # parity tests inject the N(0,1) draws
        self.weights_version = 0      # bumped by every optimizer step / load: invalidates the packed first-layer weight copies
        import os
        self.group_wgrads = os.environ.get("GO1_GROUP_WGRADS", "1") != "0"          # equal-shape wgrads of the three MLPs as grouped products
        self.fuse_tail_bwd = os.environ.get("GO1_FUSE_TAIL_BWD", "0") != "0"        # first half of the bodies' backward tails in one launch (go1_mlp_tail_backward_grouped)
        self.fuse_bias_grad = os.environ.get("GO1_FUSE_BIAS_GRAD", "1") != "0"     # bias gradients reduced in the dgrad GEMM epilogues
        self.update_streams = os.environ.get("GO1_UPDATE_STREAMS", "1") != "0"     # critic chain on a second stream during the update (measured -1.3 ms / iteration)
        self._side = None
        self.fuse_tail = os.environ.get("GO1_FUSE_TAIL", "1") != "0"     # layers behind a first layer in one tcgen05 launch (go1_mlp_tail_forward_grouped: actor + critic bodies in one grid)
        self.grads_prezeroed = False  # PPO.update zeroes the flat gradient buffer once per optimizer step (one fill instead of one per layer)

    # ------------------------------------------------------------------ flat storage
    def _ordered_params(self):
        ps = []
        for seq in (self.adaptation_module, self.actor_body, self.critic_body):
            for m in seq:
                if isinstance(m, nn.Linear):
                    ps += [m.weight, m.bias]
        return ps + [self.std]

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._flat = None              # device / dtype changed: rebuild the flat views lazily
        return r

    def flatten(self):
        """(Re)build the flat parameter/gradient buffers and re-point every parameter at its slice."""
        if self._flat is not None:          # _apply() (device/dtype moves) resets it to None
            return
        ps = self._ordered_params()
        dev = ps[0].device
        # every tensor starts on a 16-byte boundary (zero padding in between: zero gradient, never moves) so that weights
        # are TMA-readable in place wherever their row length allows it
        # The first HEAD floats of both buffers belong to no parameter: in the gradient buffer they carry the loss scalars of the
        # minibatch (KL, surrogate / value loss, ... and the adaptation MSE pair), so that ONE all-reduce per optimizer step moves
        # gradients and scalars together; [HEAD : n_adapt_params] is the adaptation module (a prefix, so its own optimizer step
        # all-reduces the prefix [0 : n_adapt_params]: scalars + adaptation gradients).
        offsets, off = {}, self.HEAD
        for p in ps:
            off = (off + 3) // 4 * 4
            offsets[id(p)] = off
            off += p.numel()
        total = (off + 3) // 4 * 4
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        for p in ps:
            o, n = offsets[id(p)], p.numel()
            flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = flat[o:o + n].view(p.shape)
        self._flat = flat
        self._packed_version = -1             # new _Net objects below: their packed weight copies do not exist yet
        self._grad = torch.zeros_like(flat)
        self.n_params = total                 # length of the flat buffers (HEAD + 3,054,619 parameters + alignment padding)
        self._nets = {}
        for name, seq in (("adapt", self.adaptation_module), ("actor", self.actor_body), ("critic", self.critic_body)):
            self._nets[name] = _Net(seq, self._flat, self._grad, offsets, self)
        self.n_adapt_params = (self._nets["adapt"].end + 3) // 4 * 4
        self.std_offset = offsets[id(self.std)]

    def ensure_packed(self):
        """Bring every packed weight copy (fused first-layer block, TMA-readable first-layer copies) up to date with the current weights.
        Called before a captured forward pass is replayed; a no-op while the weights are unchanged."""
        if self._flat is None or self._packed_version == self.weights_version:
            return
        for net in self._nets.values():
            net.refresh_packed()
        self._packed_version = self.weights_version

    @property
    def flat_params(self):
        self.flatten()
        return self._flat

    @property
    def flat_grads(self):
        self.flatten()
        return self._grad

    def load_state_dict(self, *a, **k):
        self.flatten()
        self.weights_version += 1
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------ reference API
    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        return self.std.detach().unsqueeze(0).expand_as(self._mean)

    @property
    def entropy(self):
        return (0.5 + 0.5 * torch.log(torch.tensor(2 * torch.pi)) + torch.log(self.std.detach())).sum().expand(self._mean.shape[0])

    def _impl(self):
        return int(AC_Args.gemm_impl)

    def _check_input(self, h):
        if not h.is_cuda:
            raise capi.Go1Error("ActorCritic runs on CUDA kernels only (no CPU fallback)")
        assert h.dtype == torch.float32 and h.stride(1) == 1

    def update_distribution(self, observation_history, tag="act"):
        self.flatten()
        self._check_input(observation_history)
        h = observation_history
        M, K0 = h.shape[0], self.num_obs_history
        self._a_out = self._nets["adapt"].forward(h, h.stride(0), K0, None, M, self._impl(), tag)
        latent = self._a_out[-1]
        self._p_out = self._nets["actor"].forward(h, h.stride(0), K0, latent, M, self._impl(), tag)
        self._mean = self._p_out[-1]
        self._latent = latent

    def forward_all(self, observation_history, privileged_observations, tag="act"):
        """update_distribution + evaluate in one pass.  With the tensor-core path the first layers of the three MLPs --
        which all read obs_history -- run as ONE product [M][256+512+512] = h Wcat^T: bias + ELU (+ the critic's two
        privileged columns) ride in its epilogue for the adaptation/critic slices; the actor slice is finished
        (latent columns + ELU) by go1_mlp_extra_forward once the adaptation module has produced the latent."""
        self.flatten()
        self._check_input(observation_history)
        h, priv = observation_history, privileged_observations.contiguous()
        M, K0, impl = h.shape[0], self.num_obs_history, self._impl()
        nets = self._nets
        na, npol, ncr = nets["adapt"], nets["actor"], nets["critic"]
        E = self.num_privileged_obs
        fused = impl == 1 and _Net._tma_ok(h, h.stride(0)) and K0 % 4 == 0 and 1 <= E <= 4 and \
            npol.specs[0][3] == K0 + E and ncr.specs[0][3] == K0 + E and na.specs[0][3] == K0
        if not fused:
            self.update_distribution(h, tag)
            return self._mean, self.evaluate(h, priv, tag)
        oa, op, oc = na.specs[0][2], npol.specs[0][2], ncr.specs[0][2]
        flat = self._flat

        def w1(net):
            wo, bo, o, i = net.specs[0]
            return flat[wo:wo + o * i].view(o, i), flat[bo:bo + o]

        (Wa, ba), (Wp, bp), (Wc, bc) = w1(na), w1(npol), w1(ncr)

        KP = (K0 + 31) // 32 * 32      # 128-byte row pitch of the packed weights (aligned TMA box rows, see RolloutStorage.hist_pitch)

        def build_all(old):     # the packed block [adapt | critic | actor] x K0 (128-byte row pitch), its bias row and the trailing-input
            if old is None:     # weights of the leading (adaptation | critic) columns (zeros | Wc[:, K0:]): ONE launch for all seven pieces
                old = (_empty(oa + oc + op, KP, device=flat.device)[:, :K0], _empty(1, oa + oc + op, device=flat.device),
                       _empty(oa + oc, E, device=flat.device).zero_())
            W, b, x = old
            capi.copy_segments([(W[:oa], Wa), (W[oa:oa + oc], Wc[:, :K0]), (W[oa + oc:], Wp[:, :K0]),
                                (b[:, :oa], ba.view(1, -1)), (b[:, oa:oa + oc], bc.view(1, -1)), (b[:, oa + oc:], bp.view(1, -1)), (x[oa:], Wc[:, K0:])])
            return old

        Wcat, bcat, xcat = na._cached(("l1cat", "all"), build_all)
        y = na._buf((tag, "y1cat"), M, oa + oc + op)
        na._gemm(0, 1, M, oa + oc + op, K0, h, h.stride(0), Wcat, Wcat.stride(0), y, y.stride(0), bcat, 1, 0, 1,
                 extra=priv, w_extra=xcat.data_ptr(), ld_w_extra=E, lead_cols=oa + oc)
        ya, yc, yp = y[:, :oa], y[:, oa:oa + oc], y[:, oa + oc:]
        pair = self.fuse_tail and npol._tail_ok(yp, yp.stride(0), M) and ncr._tail_ok(yc, yc.stride(0), M) and \
            [sp[2:] for sp in npol.specs[1:-1]] == [sp[2:] for sp in ncr.specs[1:-1]] and npol.specs[-1][2] + ncr.specs[-1][2] <= 16
        side = None if pair else self._side_stream(M)
        if side is not None:        # the critic's tail does not depend on the adaptation module: it runs beside adapt -> actor
            self._fork(side)
            with torch.cuda.stream(side):
                self._c_out = ncr.forward(h, h.stride(0), K0, priv, M, impl, tag, first_out=yc)
        self._a_out = na.forward(h, h.stride(0), K0, None, M, impl, tag, first_out=ya)
        latent = self._latent = self._a_out[-1]
        capi.check(capi.lib().go1_mlp_extra_forward(capi.ptr(yp), yp.stride(0), capi.ptr(latent), latent.stride(0), Wp.data_ptr() + 4 * K0, K0 + E,
                                                    M, op, E, 1, capi.stream_ptr()), "go1_mlp_extra_forward")
        if pair:                    # the equal-shape tails of the actor and critic bodies in ONE grid
            qp, outs_p, shape = npol._tail_problem(yp, yp.stride(0), M, tag)
            qc, outs_c, _ = ncr._tail_problem(yc, yc.stride(0), M, tag)
            arr = (capi.Go1TailProblem * 2)(qp, qc)
            capi.check(capi.lib().go1_mlp_tail_forward_grouped(arr, 2, M, shape[0], shape[1], shape[2], capi.stream_ptr()), "go1_mlp_tail_forward")
            self._p_out, self._c_out = [yp] + outs_p, [yc] + outs_c
            self._mean, self._value = self._p_out[-1], self._c_out[-1]
            return self._mean, self._value
        self._p_out = npol.forward(h, h.stride(0), K0, latent, M, impl, tag, first_out=yp)
        self._mean = self._p_out[-1]
        if side is not None:
            self._join(side)
        else:
            self._c_out = ncr.forward(h, h.stride(0), K0, priv, M, impl, tag, first_out=yc)
        self._value = self._c_out[-1]
        return self._mean, self._value

    # ------------------------------------------------------------------ two-stream update (independent sub-chains side by side)
    def _side_stream(self, M):
        """A second stream for the critic's chain during the update (M = minibatch rows), or None: the mid-size products leave SMs
        idle at their ramp-up and tail (one 128 x 128 tile per CTA), which an independent chain on another stream fills."""
        if not self.update_streams or M < 4096 or torch.cuda.is_current_stream_capturing():
            return None
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        return self._side

    def _fork(self, side):
        self._ev_fork.record()
        side.wait_event(self._ev_fork)

    def _join(self, side):
        self._ev_join.record(side)
        torch.cuda.current_stream().wait_event(self._ev_join)

    def act_and_evaluate(self, observation_history, privileged_observations):
        """PPO.act's two calls (actor_critic.act + evaluate, ppo.py:67-68) on one fused forward pass."""
        self.forward_all(observation_history, privileged_observations, "act")
        return self._sample(observation_history), self._value

    # Normal-like accessors used through `self.distribution`
    @property
    def mean(self):
        return self._mean

    @property
    def stddev(self):
        return self.action_std

    def act(self, observation_history, **kwargs):
        self.update_distribution(observation_history)
        return self._sample(observation_history)

    def _sample(self, observation_history):
        M = observation_history.shape[0]
        actions = torch.empty(M, self.num_actions, device=observation_history.device)
        self._logp = torch.empty(M, device=observation_history.device)
        eps = self.injected_eps
        if self._counter_dev is None or self._counter_dev.device != observation_history.device:
            with torch.inference_mode(False):
                self._counter_dev = torch.zeros(1, dtype=torch.int64, device=observation_history.device)
        capi.check(capi.lib().go1_ppo_sample_actions(capi.ptr(self._mean), self._mean.stride(0), capi.ptr(self.std.data),
                                                     capi.ptr(eps) if eps is not None else None, self.sample_seed, 0, capi.ptr(self._counter_dev),
                                                     capi.ptr(actions), capi.ptr(self._logp), M, self.num_actions, capi.stream_ptr()), "sample")
        self._last_actions = actions
        return actions

    def get_actions_log_prob(self, actions):
        last = getattr(self, "_last_actions", None)
        if last is not None and actions.data_ptr() == last.data_ptr() and actions.shape == last.shape:      # the sample kernel already produced it
            return self._logp
        d = actions - self._mean
        sd = self.std.detach()
        return (-(d * d) / (2 * sd * sd) - torch.log(sd) - 0.9189385332046727).sum(-1)

    def act_expert(self, ob, policy_info={}):
        return self.act_teacher(ob["obs_history"], ob["privileged_obs"])

    def act_inference(self, ob, policy_info={}):
        return self.act_student(ob["obs_history"], policy_info=policy_info)

    def act_student(self, observation_history, policy_info={}):
        if observation_history.shape[0] == 0:
            return observation_history.new_zeros(0, self.num_actions)
        self.update_distribution(observation_history, tag="student")
        policy_info["latents"] = self._latent.detach().cpu().numpy()
        return self._mean

    def act_teacher(self, observation_history, privileged_info, policy_info={}):
        if observation_history.shape[0] == 0:
            return observation_history.new_zeros(0, self.num_actions)
        self.flatten()
        h = observation_history
        out = self._nets["actor"].forward(h, h.stride(0), self.num_obs_history, privileged_info.contiguous(), h.shape[0], self._impl(), "teacher")
        policy_info["latents"] = privileged_info
        return out[-1]

    def evaluate(self, observation_history, privileged_observations, tag="act", **kwargs):
        self.flatten()
        self._check_input(observation_history)
        h = observation_history
        self._c_out = self._nets["critic"].forward(h, h.stride(0), self.num_obs_history, privileged_observations.contiguous(), h.shape[0], self._impl(), tag)
        self._value = self._c_out[-1]
        return self._value

    def get_student_latent(self, observation_history):
        self.flatten()
        h = observation_history
        return self._nets["adapt"].forward(h, h.stride(0), self.num_obs_history, None, h.shape[0], self._impl(), "latent")[-1]

    # ------------------------------------------------------------------ explicit backward passes (ppo.py:154-189)
    def _flush_wgrads(self):
        """Launch the queued wgrads: equal shapes (same M, N, K and operand strides) as ONE grouped product (go1_gemm_grouped: up to four
        problems in one grid), the rest one by one.  All of them accumulate into the pre-zeroed flat gradient buffer."""
        import ctypes as C
        q, self._wgrad_queue = self._wgrad_queue, None
        groups = {}
        for it in q:
            groups.setdefault(it[:6], []).append(it)
        L, st = capi.lib(), capi.stream_ptr()
        for (o, K, M, ldz, ld_in, i), items in groups.items():
            for k0 in range(0, len(items), 4):
                chunk = items[k0:k0 + 4]
                n = len(chunk)
                A = (C.c_void_p * n)(*[it[6].data_ptr() for it in chunk])
                B = (C.c_void_p * n)(*[it[7].data_ptr() for it in chunk])
                Cc = (C.c_void_p * n)(*[it[8].data_ptr() for it in chunk])
                capi.check(L.go1_gemm_grouped(1, 0, o, K, M, n, A, ldz, B, ld_in, Cc, i, 1, st), "go1_gemm_grouped")

    def backward_ppo(self, h, priv, dmean, dvalue, dstd, aug=False):
        """Gradients of the PPO loss into flat_grads (overwrites). h/priv are the minibatch inputs of the forward
        pass just run with tag='train'; dmean [M,A], dvalue [M,1], dstd [A].
        aug: h is a view of a row buffer with spare columns behind the K0 history columns that the caller has filled with
        [1 | priv (E) | anything (E)] (RolloutStorage.mini_batch_generator does).  The latent is copied into the last E and the fused
        first-layer wgrad runs over K0 + 1 + 2E input columns: its extra output columns ARE the three first-layer bias gradients and the
        trailing-input weight gradients of the critic (priv columns) and the actor (latent columns), for free on the tensor core, so the
        dgrad epilogues that produce the first-layer dz skip those reductions."""
        M, K0, impl = h.shape[0], self.num_obs_history, self._impl()
        nets = self._nets
        if impl == 1 and M >= 64 and _Net._tma_ok(h, h.stride(0)):
            # the three first layers share their input: ONE dz [M][o_a+o_p+o_c] and ONE tensor-core wgrad
            # dWcat[1280][2100] = dzcat^T h instead of three (the first-layer dz of each net is written straight into its
            # column slice of `dz1` by that net's layer-2 dgrad)
            oa, op, oc = nets["adapt"].specs[0][2], nets["actor"].specs[0][2], nets["critic"].specs[0][2]
            dz1 = nets["adapt"]._buf(("train", "dz1cat"), M, oa + op + oc)
            E = self.num_privileged_obs
            aug = bool(aug) and self.fuse_bias_grad and self.grads_prezeroed and h.stride(0) >= K0 + 1 + 2 * E and 1 <= E <= 4 and \
                nets["actor"].specs[0][3] == K0 + E and nets["critic"].specs[0][3] == K0 + E and priv.shape[1] == E
            KA = K0 + 1 + 2 * E if aug else K0          # input columns of the fused wgrad
            if aug:
                h_ext = h.as_strided((M, KA), (h.stride(0), 1))
                capi.copy_segments([(h_ext[:, K0 + 1 + E:], self._latent)])
            if self.group_wgrads and self.grads_prezeroed:
                self._wgrad_queue = []
            pre_p = pre_c = None
            if self.fuse_tail_bwd:      # first half of both bodies' backward tails in ONE grid (dz3, dz2 and their bias gradients)
                if nets["actor"].tail_bwd_ok(self._p_out, dmean) and nets["critic"].tail_bwd_ok(self._c_out, dvalue):
                    tp, tc = nets["actor"].tail_bwd_problem(self._p_out, dmean, M, "train"), nets["critic"].tail_bwd_problem(self._c_out, dvalue, M, "train")
                    arr = (capi.Go1TailBwdProblem * 2)(tp[0], tc[0])
                    capi.check(capi.lib().go1_mlp_tail_backward_grouped(arr, 2, M, 128, 256, capi.stream_ptr()), "go1_mlp_tail_backward")
                    pre_p, pre_c = tp[1:], tc[1:]
            side = self._side_stream(M)
            if side is not None:    # critic chain beside actor -> adaptation chain
                self._fork(side)
                with torch.cuda.stream(side):
                    nets["critic"].backward(h, h.stride(0), K0, priv, self._c_out, dvalue, M, impl, 0, tag="train", dz1_out=dz1[:, oa + op:], aug_first=aug, pre=pre_c)
            dlat = nets["actor"].backward(h, h.stride(0), K0, self._latent, self._p_out, dmean, M, impl, 0, want_dextra=True, tag="train", dz1_out=dz1[:, oa:oa + op],
                                          aug_first=aug, pre=pre_p)
            if side is None:
                nets["critic"].backward(h, h.stride(0), K0, priv, self._c_out, dvalue, M, impl, 0, tag="train", dz1_out=dz1[:, oa + op:], aug_first=aug, pre=pre_c)
            nets["adapt"].backward(h, h.stride(0), K0, None, self._a_out, dlat, M, impl, 0, tag="train", dz1_out=dz1[:, :oa], aug_first=aug)
            if side is not None:
                self._join(side)
            if self._wgrad_queue is not None:
                self._flush_wgrads()
            n0 = nets["adapt"]
            KP = (KA + 31) // 32 * 32
            gcat = n0._buf(("train", "gWcat"), oa + op + oc, KP)
            n0._gemm(1, 0, oa + op + oc, KA, M, dz1, dz1.stride(0), h_ext if aug else h, h.stride(0), gcat, KP, None, 0, 0, 1)
            row, pairs = 0, []
            for name, xcol in (("adapt", None), ("actor", K0 + 1 + E), ("critic", K0 + 1)):
                wo, bo, o, i = nets[name].specs[0]
                gW = self._grad[wo:wo + o * i].view(o, i)
                pairs.append((gW[:, :K0], gcat[row:row + o, :K0]))
                if aug:     # column K0: the bias gradient; columns xcol..xcol+E: the trailing-input weight gradient of this net
                    pairs.append((self._grad[bo:bo + o].view(o, 1), gcat[row:row + o, K0:K0 + 1]))
                    if xcol is not None:
                        pairs.append((gW[:, K0:K0 + E], gcat[row:row + o, xcol:xcol + E]))
                row += o
            capi.copy_segments(pairs)
        else:
            dlat = nets["actor"].backward(h, h.stride(0), K0, self._latent, self._p_out, dmean, M, impl, 0, want_dextra=True, tag="train")
            nets["critic"].backward(h, h.stride(0), K0, priv, self._c_out, dvalue, M, impl, 0, tag="train")
            nets["adapt"].backward(h, h.stride(0), K0, None, self._a_out, dlat, M, impl, 0, tag="train")
        self._grad[self.std_offset:self.std_offset + self.num_actions].copy_(dstd)

    def backward_adaptation(self, h, outs, dpred):
        M, K0 = h.shape[0], self.num_obs_history
        self._nets["adapt"].backward(h, h.stride(0), K0, None, outs, dpred, M, self._impl(), 0, tag="adapt")

    def adaptation_forward(self, h):
        self.flatten()
        return self._nets["adapt"].forward(h, h.stride(0), self.num_obs_history, None, h.shape[0], self._impl(), "adapt")


def get_activation(act_name):
    table = {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "crelu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    if act_name not in table:
        print("invalid activation function!")
        return None
    return table[act_name]()
