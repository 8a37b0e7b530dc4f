"""Runs every flash-attention test case without stopping at the first failure and prints per-tensor errors
(one gpurun call shows the whole picture).  ``python benchmarks/debug_flash_attn.py [variant]``"""

import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main() -> None:
    from test_flash_attn_gpu import CASES, _oracle

    from d9d_b200 import ops as _ops
    from d9d_b200.kernel.flash_attn import flash_attn_func

    _ops.load()
    for variant in ([int(sys.argv[1])] if len(sys.argv) > 1 else [0]):
        os.environ["D9D_FA_VARIANT"] = str(variant)
        for case in CASES:
            B, Sq, Sk, Hq, Hk, D, causal, window, use_sink, softcap = case
            torch.manual_seed(Sq * 7 + D + Hq)
            q = torch.randn(B, Sq, Hq, D, device="cuda").bfloat16().requires_grad_()
            k = torch.randn(B, Sk, Hk, D, device="cuda").bfloat16().requires_grad_()
            v = torch.randn(B, Sk, Hk, D, device="cuda").bfloat16().requires_grad_()
            sink = torch.randn(Hq, device="cuda").requires_grad_() if use_sink else None
            dout = torch.randn(B, Sq, Hq, D, device="cuda").bfloat16()
            line = f"variant={variant} case={case}: "
            try:
                out, lse = flash_attn_func(q, k, v, causal=causal, window_size=window, learnable_sink=sink, softcap=softcap,
                                           return_lse=True)
                torch.cuda.synchronize()
                ref_out, ref_lse, (rdq, rdk, rdv), rdsink = _oracle(q, k, v, causal, window, sink, softcap, dout)

                def err(a, b):
                    a, b = a.float(), b.float()
                    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-6)

                fin = torch.isfinite(ref_lse)
                line += f"out {err(out, ref_out):.4f} lse {float((lse[fin] - ref_lse[fin]).abs().max()):.5f} "
                out.backward(dout)
                torch.cuda.synchronize()
                line += f"dq {err(q.grad, rdq):.4f} dk {err(k.grad, rdk):.4f} dv {err(v.grad, rdv):.4f}"
                if use_sink:
                    line += f" dsink {err(sink.grad, rdsink):.4f}"
            except Exception as exc:  # noqa: BLE001
                line += f"EXCEPTION {type(exc).__name__}: {str(exc)[:300]}"
                traceback.print_exc()
            print(line, flush=True)


if __name__ == "__main__":
    main()
