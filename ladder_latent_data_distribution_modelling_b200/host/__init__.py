"""Host-side mirror of the reference's `codes/` package (same class / function names, argument
meaning and printed lines) on top of the sm_100a engine.  `codes/` at the repo root re-exports
these modules so `from codes.models import MNISTModel_digit` keeps working."""
