"""Import-compatible stand-in for the reference's ``models`` package
(src/scripts/benchmark/models): ``from models import multimodalIntraInterModal`` resolves to
the B200 implementation when this directory's parent is placed first on sys.path."""
