"""The reference's own `tests/test_utils.py` (class `TestTransposeForDisplay`) restated for this package: same shapes
(2-D arrays included), same assertions -- shape swapped along the first two axes, multiset of values unchanged, the
algebra of transposing with and without the vertical flip.  CPU, and the same on CUDA tensors."""
import numpy as np
import pytest
import torch

from jaxrenderer_b200.utils import transpose_for_display

SHAPES = [(1, 1), (1, 7), (3, 1), (20, 31, 3), (11, 11, 4)]


def _matrix(shape, device="cpu"):
    g = torch.Generator().manual_seed(20230701)
    return torch.rand(shape, generator=g).to(device)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("flip_vertical", [True, False])
def test_transposed_shape_must_be_flipped_along_first_two_axis(shape, flip_vertical):
    matrix = _matrix(shape)
    transposed = transpose_for_display(matrix, flip_vertical=flip_vertical)
    assert tuple(matrix.shape) == tuple(shape), "Matrix shape must not be changed"
    assert tuple(transposed.shape) == (shape[1], shape[0], *shape[2:]), "Transposed shape must be flipped along first two axises"


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("flip_vertical", [True, False])
def test_transposed_unique_values_and_count_must_be_the_same(shape, flip_vertical):
    matrix = _matrix(shape)
    transposed = transpose_for_display(matrix, flip_vertical=flip_vertical)
    m, m_cnt = np.unique(matrix.numpy(), return_counts=True)
    t, t_cnt = np.unique(transposed.numpy(), return_counts=True)
    assert (m == t).all(), "Unique values must be the same"
    assert (m_cnt == t_cnt).all(), "Unique values count must be the same"


def _flip_algebra(matrix):
    tf_f = lambda x: transpose_for_display(x, flip_vertical=True)    # noqa: E731
    t_f = lambda x: transpose_for_display(x, flip_vertical=False)    # noqa: E731
    tf, t = tf_f(matrix), t_f(matrix)
    assert bool((t != tf).any()), "flipped vertical will change the matrix"
    assert torch.equal(t_f(tf_f(t_f(tf_f(matrix)))), matrix), "flip twice, transpose 4 times should be identity"
    assert torch.equal(t_f(t_f(matrix)), matrix), "transpose twice should be identity"
    assert torch.equal(tf_f(tf_f(tf_f(tf_f(matrix)))), matrix), "transpose and flip 4 times should be identity"
    assert bool((tf_f(tf_f(matrix)) != matrix).any()), "transpose and flip twice should not be identity"


@pytest.mark.parametrize("shape", [(5, 3), (20, 31, 3), (11, 11, 4)])
def test_flip_vertical(shape):
    _flip_algebra(_matrix(shape))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 3), (20, 31, 3), (11, 11, 4)])
def test_flip_vertical_on_cuda_tensors(shape):
    m = _matrix(shape, "cuda")
    _flip_algebra(m)
    assert torch.equal(transpose_for_display(m).cpu(), transpose_for_display(m.cpu()))
