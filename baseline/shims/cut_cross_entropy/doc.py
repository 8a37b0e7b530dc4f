LINEAR_CROSS_ENTROPY_DOC = "Linear cross-entropy: loss(e @ c.T + bias, targets) without materialising all logits."
CCE_OPTS_DOC = [
    ":param filter_eps: gradient filtering threshold (ignored by this stand-in: gradients are exact).",
    ":param accum_e_fp32: accumulate the embedding gradient in fp32.",
    ":param accum_c_fp32: accumulate the classifier gradient in fp32.",
    ":param filter_e_grad: filter the embedding gradient.",
    ":param filter_c_grad: filter the classifier gradient.",
]
IMPL_DOC = ":param impl: implementation name."
DTENSOR_NOTE = "DTensor classifiers are gathered to full tensors."


def _append(fn, text: str, start: bool):
    doc = fn.__doc__ or ""
    fn.__doc__ = (text + "\n" + doc) if start else (doc + "\n" + text)
    return fn


def add_doc_start(*texts: str):
    def deco(fn):
        return _append(fn, "\n".join(texts), True)

    return deco


def add_doc_end(*texts: str):
    def deco(fn):
        return _append(fn, "\n".join(texts), False)

    return deco
