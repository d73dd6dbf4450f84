"""``ending_sentence`` hook kept for call-compatibility with the reference's
``qsextra/tools/oracle.py:26-28`` (prints a closing line when ``verbose``).

The module name is the reference's (its "oracle" prints a sentence at the end of a run); it has nothing to do with the
repository-level test oracle ``oracle/``, which no product module imports."""


def ending_sentence(verbose: bool):
    if verbose:
        print('Done.')
