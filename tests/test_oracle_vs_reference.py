"""Live re-check of the golden fixtures against the reference (only where /root/reference exists)."""
import pytest


@pytest.mark.reference
def test_goldens_reproduce_from_reference():
    from oracle import gen_golden
    assert gen_golden.main(check=True)
