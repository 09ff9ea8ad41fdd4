import pytest

pytestmark = pytest.mark.gpu


def test_smoke():
    import __graft_entry__ as g
    g.smoke()
