"""Mirror of the field bag of reference preprocessor/radar_point_cloud.py:18-37 (the input of
``build_geometric_graph``).  The NaN / range filters of the reference are dataset plumbing and
out of scope."""


class RadarPointCloud():
    def __init__(self):
        self.X_cc = None
        self.X_seq = None

        self.V_cc = None
        self.V_cc_compensated = None

        self.range_sc = None
        self.azimuth_sc = None
        self.rcs = None

        self.vr = None
        self.vr_compensated = None

        self.timestamp = None
        self.sensor_id = None

        self.uuid = None
        self.track_id = None
        self.label_id = None
