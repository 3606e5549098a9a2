"""B200-native spectral-element operator engine (drop-in for the reference's
``sem.discrete`` / ``sem.basis_functions`` / ``sem.quadratures`` API on the
Poisson hot path).  See DESIGN.md."""
__version__ = "0.1.0"
