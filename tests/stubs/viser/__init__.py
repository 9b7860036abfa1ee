"""Test stub of the third-party ``viser`` web viewer (reference main.py:7-8,42-43,64-83,129-137 creates a server
unconditionally: ``--viewer`` is ``store_true`` with ``default=True``).  UI is out of scope;
this stub only records the calls so the UNMODIFIED reference driver can run head-less in tests/test_reference_driver.py."""
import contextlib

CALLS = []


class _Number:
    def __init__(self, name, initial_value):
        self.name, self.value = name, initial_value


class _Gui:
    @contextlib.contextmanager
    def add_folder(self, name):
        CALLS.append(("add_folder", name))
        yield self

    def add_number(self, name, initial_value=0, disabled=False):
        CALLS.append(("add_number", name))
        return _Number(name, initial_value)


class _Scene:
    def add_mesh_simple(self, name, vertices, faces, **kw):
        assert vertices.ndim == 2 and vertices.shape[1] == 3 and faces.ndim == 2 and faces.shape[1] == 3
        CALLS.append(("add_mesh_simple", name))


class ViserServer:
    def __init__(self, port=8080, **kw):
        self.port, self.gui, self.scene = port, _Gui(), _Scene()
