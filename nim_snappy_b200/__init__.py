"""Import shim: the package directory is `nim-snappy_b200/` (not a valid Python identifier), so
`import nim_snappy_b200` resolves here and re-roots itself onto that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nim-snappy_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
