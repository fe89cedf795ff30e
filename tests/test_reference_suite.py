"""The reference's OWN test files against this package (host side of the drop-in).  Only where /root/reference exists
(the build container); oracle/run_reference_tests.sh is the recipe, oracle/reference_tests_shim.py explains what is
substituted (the package name, the absent optional dependencies, the device plan -> oracle-backed stand-in).

Every reference test either passes or is on the list below with the reason it cannot; a test that starts failing, or
a listed one that starts passing, fails this test so that the list stays true."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXPECTED_FAILURES = {
    # astropy WCS / RA-DEC placement and its serialised state: out of scope
    "test_window.py::TestWindow::test_window_creation": "get_astropywcs",
    "test_window.py::TestWindow::test_window_errors": "reference_radec",
    "test_window.py::TestWindow::test_window_state": "projection in the state",
    "test_window_list.py::TestWindowList::test_image_list_errors": "ConflicingWCS",
    "test_image_list.py::TestImageList::test_image_list_errors": "ConflicingWCS",
    "test_image.py::TestImage::test_image_wcs_roundtrip": "astropy WCS",
    "test_image_header.py::TestImageHeader::test_iamge_header_repr": "WCS state",
    "test_image_header.py::TestImageHeader::test_image_header_creation": "Image_Header without data",
    "test_image_header.py::TestImageHeader::test_image_header_wcs_roundtrip": "astropy WCS",
    "test_image.py::TestImage::test_image_save_load": "FITS save / load",
    "test_image.py::TestTargetImage::test_target_save_load": "FITS save / load",
    # peripheral image methods: physical-unit expand(), Lanczos shift_origin(); and the reference's 4-element
    # Image.crop, which cuts the data as [x lo, x hi, y lo, y hi] but its window as [x lo, y lo, x hi, y hi]
    "test_image.py::TestImage::test_image_errors": "Image.expand",
    "test_image.py::TestModelImage::test_shift": "Model_Image.shift_origin",
    "test_image.py::TestImage::test_image_manipulation": "4-element crop convention",
    "test_group_models.py::TestPSFGroup::test_psfgroupmodel_saveload": "psf group model",
    # CPU torch convolution / interpolation helpers: deliberately absent (convolution lives on the device only),
    # star-stacking PSF construction, hdf5
    "test_utils.py::TestFFT::test_fft": "no CPU convolution path",
    "test_utils.py::TestFFT::test_fft_multi": "no CPU convolution path",
    "test_utils.py::TestInterpolate::test_interpolate_functions": "utils.interpolate",
    "test_utils.py::TestPSF::test_make_psf": "construct_psf",
    "test_utils.py::TestConversions::test_conversion_dict_to_hdf5": "h5py",
    # model families outside north_star, PSF group models
    "test_group_models.py::TestPSFGroup::test_psfgroupmodel_creation": "psf group model",
    "test_group_models.py::TestPSFGroup::test_psfgroupmodel_fitting": "psf group model",
    "test_psfmodel.py::TestEigenPSF::test_init": "eigen psf model",
    "test_psfmodel.py::TestPixelPSF::test_init": "pixelated psf model",
    # optimisers other than LM / Iter / Iter_LM: out of scope
    "test_fit.py::TestComponentModelFits::test_sersic_fit_grad": "fit.Grad",
    "test_fit.py::TestGroupModelFits::test_groupmodel_fit": "fit.Grad",
    "test_fit.py::TestMiniFit::test_minifit": "fit.MiniFit",
    "test_fit.py::TestHMC::test_hmc_sample": "fit.HMC",
    "test_fit.py::TestNUTS::test_nuts_sample": "fit.NUTS",
    "test_fit.py::TestMHMCMC::test_singlesersic": "fit.MHMCMC",
    # image_chunksize=15 on a 50-pixel window: the reference's chunk rounding leaves its Jacobian 2 pixels short;
    # lowering refuses that instead of reproducing it
    "test_fit.py::TestLM::test_chunk_image_jacobian": "chunk rounding",
}


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="the reference only exists in the build container")
def test_reference_test_files_against_this_package():
    out = subprocess.run(["bash", os.path.join(ROOT, "oracle", "run_reference_tests.sh")], capture_output=True, text=True,
                         timeout=900).stdout
    passed = {t.rsplit("/", 1)[-1] for t in re.findall(r"^PASSED (\S+)", out, flags=re.M)}
    failed = {t.rsplit("/", 1)[-1] for t in re.findall(r"^(?:FAILED|ERROR) (\S+)", out, flags=re.M)}
    assert failed == set(EXPECTED_FAILURES), (sorted(failed - set(EXPECTED_FAILURES)), sorted(set(EXPECTED_FAILURES) - failed))
    assert len(passed) >= 78, out[-2000:]
