from .reductions import init_reductions

init_reductions()
