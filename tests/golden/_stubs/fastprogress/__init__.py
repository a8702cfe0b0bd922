"""Minimal stand-in for the `fastprogress` package (absent from this image, no network).

Only used by tests/golden/make_golden.py so that the *reference* (`/root/reference/littlemcmc`)
can be imported in this container.  The reference touches exactly
`fastprogress.fastprogress.progress_bar(iterable, total=None, display=True)` with a `.comment`
attribute and an `.update()` method (reference littlemcmc/sampling.py:23,457-459;
parallel_sampling.py:26,443-445,460,471).
"""
