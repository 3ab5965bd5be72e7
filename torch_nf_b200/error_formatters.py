"""Error-message helpers with the reference's behaviour
(reference torch_nf/error_formatters.py:4-34)."""

def format_type_err_msg(obj, arg_name, arg, correct_type):
    """"<Class> argument <name> must be <type> not <type>." -- raises ValueError
    when the argument already has the correct type (error_formatters.py:16-18)."""
    got = arg.__class__
    if got is correct_type:
        raise ValueError("Invalid TypeError message: type(arg) == correct_type.")
    return "%s argument %s must be %s not %s." % (
        obj.__class__.__name__, arg_name, correct_type.__name__, got.__name__)


def dbg_check(tensor, name):
    """Print inf / nan counts of a tensor and return a truthy value if any
    (error_formatters.py:26-34)."""
    import torch
    total = 1
    for n in tensor.shape:
        total *= n
    n_inf = int(torch.isinf(tensor).sum().item())
    n_nan = int(torch.isnan(tensor).sum().item())
    print(name, "infs %d/%d" % (n_inf, total), "nans %d/%d" % (n_nan, total))
    return n_nan or n_inf
