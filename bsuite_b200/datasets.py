"""MNIST training set loader in the reference's on-disk format.

The reference (`bsuite/utils/datasets.py:42-69`) downloads four idx-ubyte .gz
files into /tmp/mnist and parses the images as **int8** (`:55-56`), so pixels
>= 128 become negative after the `/255` in `mnist.py:64`.  This loader reads the
same files from the same default directory with the same int8 view; it never
downloads (there is no network on the GPU boxes).  `write_synthetic_mnist`
creates files of identical format from a seeded generator so the mnist families
can run (and be parity-tested against the reference) offline.
"""

import gzip
import os
import struct
from typing import Optional, Tuple

import numpy as np

DEFAULT_DIR = '/tmp/mnist'
ENV_VAR = 'BSUITE_B200_MNIST_DIR'
TRAIN_IMAGES = 'train-images-idx3-ubyte.gz'
TRAIN_LABELS = 'train-labels-idx1-ubyte.gz'
TEST_IMAGES = 't10k-images-idx3-ubyte.gz'
TEST_LABELS = 't10k-labels-idx1-ubyte.gz'


def _read_images(path: str) -> np.ndarray:
  with gzip.open(path, 'rb') as fh:
    _, count, rows, cols = struct.unpack('>IIII', fh.read(16))
    raw = np.frombuffer(fh.read(), dtype=np.uint8)
  return raw.view(np.int8).reshape(count, rows, cols)    # int8 reinterpretation, as the reference


def _read_labels(path: str) -> np.ndarray:
  with gzip.open(path, 'rb') as fh:
    fh.read(8)
    return np.frombuffer(fh.read(), dtype=np.uint8).copy()


def resolve_dir(data_dir: Optional[str] = None) -> str:
  return data_dir or os.environ.get(ENV_VAR) or DEFAULT_DIR


def load_mnist_train(data_dir: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
  """Returns (images int8 [n,28,28], labels uint8 [n]) of the training split."""
  directory = resolve_dir(data_dir)
  images_path = os.path.join(directory, TRAIN_IMAGES)
  labels_path = os.path.join(directory, TRAIN_LABELS)
  if not (os.path.isfile(images_path) and os.path.isfile(labels_path)):
    raise FileNotFoundError(
        f'MNIST idx files not found in {directory!r}. Place {TRAIN_IMAGES} and {TRAIN_LABELS} there (or set '
        f'${ENV_VAR}); for offline use call bsuite_b200.datasets.write_synthetic_mnist(directory).')
  return _read_images(images_path), _read_labels(labels_path)


def write_synthetic_mnist(directory: str, num_train: int = 2048, num_test: int = 256, seed: int = 0) -> str:
  """Writes seeded random idx-ubyte .gz files (all four names the reference expects)."""
  os.makedirs(directory, exist_ok=True)
  rng = np.random.RandomState(seed)
  for images_name, labels_name, count in ((TRAIN_IMAGES, TRAIN_LABELS, num_train), (TEST_IMAGES, TEST_LABELS, num_test)):
    pixels = rng.randint(0, 256, size=(count, 28, 28)).astype(np.uint8)
    labels = rng.randint(0, 10, size=count).astype(np.uint8)
    with gzip.open(os.path.join(directory, images_name), 'wb') as fh:
      fh.write(struct.pack('>IIII', 2051, count, 28, 28))
      fh.write(pixels.tobytes())
    with gzip.open(os.path.join(directory, labels_name), 'wb') as fh:
      fh.write(struct.pack('>II', 2049, count))
      fh.write(labels.tobytes())
  return directory
