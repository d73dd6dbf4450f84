"""``ending_sentence`` hook kept for call-compatibility with the reference's
``qsextra/tools/oracle.py:26-28`` (prints a closing line when ``verbose``)."""


def ending_sentence(verbose: bool):
    if verbose:
        print('Done.')
