"""Stub of the un-vendored third-party `kanga` package (eeyore/chains/chain_list.py:8 imports
kanga.chains.ChainArray at module load).  It contributes no arithmetic to the hot path; this stub only
lets oracle/make_golden.py import the unmodified reference in the build container."""
