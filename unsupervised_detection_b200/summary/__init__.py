"""TensorBoard event-file output (the reference's tf.summary.* + Supervisor summary_writer) without TensorFlow."""
from .writer import SummaryWriter, read_events, normalize_float_image, histogram_proto  # noqa: F401
