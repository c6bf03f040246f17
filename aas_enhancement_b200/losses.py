"""``L1Loss_mask`` -- the loss every trainer of the reference applies to the LMFB features right
after the front-end (Speech_enhancement_by_AAS/model.py:19-31; call sites trainer_AAS.py:146-161,
:176-181, trainer_DCE.py, trainer_FSEGAN.py).  Same name, call signature and return value
``(loss, nElement)``; the sum and its gradient run in CUDA kernels behind the C ABI
(``aas_l1_abs_sum`` / ``aas_l1_abs_grad``), deterministically.

Reference quirks, reproduced by default (``fix_masking=False``):
* ``err.masked_fill(mask, 0)`` is not in-place, i.e. a no-op: padded frames contribute to the sum;
* the divisor is ``mask.nelement() - mask.sum()``: the number of unmasked FRAMES, not elements;
* ``nElement`` is only defined when ``mask[0][0][0] == 0`` (the reference raises otherwise).
``fix_masking=True`` zeroes the padded frames, which is the evident intent.
"""
from __future__ import annotations

import torch

from . import _lib


class _L1AbsSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, target, mask_or_none, rows=None):
        if not input.is_cuda:
            raise RuntimeError("L1Loss_mask is CUDA-only (sm_100a); there is no CPU fallback")
        lib = _lib.load()
        a = input.contiguous().float()
        b = target.contiguous().float()
        if a.shape != b.shape or a.dim() != 3:
            raise ValueError("input and target must both be (N, C, Tmax)")
        n, c, tmax = a.shape
        m = None
        if mask_or_none is not None:
            m = mask_or_none.contiguous().to(torch.uint8)
            if m.shape != (n, 1, tmax):
                raise ValueError("mask must be (N, 1, Tmax)")
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            stream = torch.cuda.current_stream(a.device).cuda_stream
            if rows is not None and m is None:
                # the per-row sums come from the front-end's CMVN kernel (LMFBFrontEnd(..., l1_target=target)):
                # only the (N*C)-term final sum is left, `input` is not read again
                r = rows.contiguous().float()
                if r.numel() != n * c:
                    raise ValueError("rows must hold one partial sum per (utterance, channel) row")
                rc = lib.aas_l1_rows_sum(r.data_ptr(), n * c, out.data_ptr(), stream)
            else:
                partial = torch.empty(lib.aas_l1_partial_count(), dtype=torch.float32, device=a.device)
                rc = lib.aas_l1_abs_sum(a.data_ptr(), b.data_ptr(), None if m is None else m.data_ptr(),
                                        n, c, tmax, partial.data_ptr(), out.data_ptr(), stream)
        _lib.check(rc)
        ctx.save_for_backward(a, b, m)
        return out[0]

    @staticmethod
    def backward(ctx, grad):
        a, b, m = ctx.saved_tensors
        lib = _lib.load()
        n, c, tmax = a.shape
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ga = torch.empty_like(a) if need_a else None
        gb = torch.empty_like(b) if need_b else None
        if ga is None and gb is None:
            return None, None, None, None
        scale = grad.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            rc = lib.aas_l1_abs_grad(a.data_ptr(), b.data_ptr(), None if m is None else m.data_ptr(),
                                     n, c, tmax, scale.data_ptr(),
                                     None if ga is None else ga.data_ptr(),
                                     None if gb is None else gb.data_ptr(),
                                     torch.cuda.current_stream(a.device).cuda_stream)
        _lib.check(rc)
        return ga, gb, None, None


class L1Loss_mask(torch.nn.Module):
    """Drop-in for the reference's ``L1Loss_mask`` (model.py:19-31): ``forward(input, target, mask)``
    returns ``(loss, nElement)``."""

    def __init__(self, fix_masking: bool = False):
        super().__init__()
        self.fix_masking = fix_masking

    def forward(self, input, target, mask, rows=None):
        """``rows``: optional per-row sums ``sum_t |input - target|`` from the front-end's epilogue
        (``LMFBFrontEnd.forward(..., l1_target=target)``); the loss value then costs one tiny kernel and
        ``input`` is not read again (not with ``fix_masking``, whose sum skips the padded frames)."""
        mask_sum = mask.sum()
        if bool(mask[0][0][0] == 0):                      # data_as_0 = True (model.py:25)
            n_element = mask.nelement() - mask_sum
        else:                                             # the reference hits an UnboundLocalError here
            raise RuntimeError("L1Loss_mask: mask[0][0][0] != 0 -- nElement is undefined in the reference "
                               "(model.py:25-26); batches are length-sorted, so the first frame is never padding")
        err_sum = _L1AbsSum.apply(input, target, mask if self.fix_masking else None,
                                  None if self.fix_masking else rows)
        loss = err_sum / n_element
        return loss, n_element


class CTCLoss(torch.nn.Module):
    """Stand-in for ``warpctc_pytorch.CTCLoss`` (un-vendored; imported at trainer_AAS.py:12 and
    called at :168 as ``self.CTCLoss(prob, targets, sizes, target_sizes) / N``) -- SURVEY 8(f)
    rank 4.  Same call signature and semantics: ``acts`` are UNNORMALISED activations
    ``(T, N, C)`` (warp-ctc applies the softmax itself), ``labels`` the flat int32 concatenation
    of the targets (blank = 0), ``act_lens`` / ``label_lens`` int32 on the host; the result is the
    cost SUMMED over the batch as a 1-element tensor (the trainer divides by N).  Library code
    (``torch.nn.functional.ctc_loss``) -- it is the boundary after the front-end, not part of the
    hand-written path."""

    def __init__(self, size_average: bool = False, length_average: bool = False):
        super().__init__()
        self.size_average, self.length_average = size_average, length_average

    def forward(self, acts, labels, act_lens, label_lens):
        log_probs = torch.nn.functional.log_softmax(acts.float(), dim=2)
        cost = torch.nn.functional.ctc_loss(log_probs, labels.to(torch.long), act_lens.to(torch.long),
                                            label_lens.to(torch.long), blank=0, reduction="sum",
                                            zero_infinity=False)
        if self.length_average:
            cost = cost / act_lens.sum().to(cost.dtype)
        elif self.size_average:
            cost = cost / acts.size(1)
        return cost.reshape(1)
