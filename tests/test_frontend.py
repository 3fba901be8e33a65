"""Configuration front-end: the reference's command-line flags
(mimo/main.cc:174-240) and the GUI's JSON device record (Interface/usrp_device.cpp:13-46)."""
import json

import pytest

import rub_mimo_b200 as rub


def test_reference_command_line_parses_unchanged():
    cfg, fe = rub.config_from_args(["mimo", "-f", "2.45e9", "--rate", "2e6", "--dsp_gain=0.5", "--tx_gain", "40",
                                    "--rx_gain=35.5", "--num_subcarriers", "1024", "--cp_len=72",
                                    "--rx_addr", "serial=30C51D4", "--tx_addr=serial=30C5426",
                                    "--tx_subdev", "A:A A:B", "--rx_subdev=A:A A:B", "-v"])
    assert (cfg.M, cfg.cp_len) == (1024, 72)
    assert fe.cent_freq == 2.45e9 and fe.samp_rate == 2e6 and fe.dsp_gain == 0.5
    assert fe.txgain == 40.0 and fe.rxgain == 35.5 and fe.verbose == 1 and fe.help == 0
    assert fe.rx_addr == b"serial=30C51D4" and fe.tx_addr == b"serial=30C5426"
    assert fe.tx_subdev == b"A:A A:B" and fe.rx_subdev == b"A:A A:B"
    cfg.validate()


def test_defaults_survive_and_quiet_and_help():
    base = rub.Config(M=2048, cp_len=152)
    cfg, fe = rub.config_from_args(["mimo", "-q", "--help"], cfg=base)
    assert (cfg.M, cfg.cp_len, cfg.nac) == (2048, 152, base.nac)
    assert fe.verbose == 0 and fe.help == 1 and abs(fe.dsp_gain - 0.25) < 1e-7


@pytest.mark.parametrize("argv", [["mimo", "--bogus"], ["mimo", "--cp_len"], ["mimo", "--num_subcarriers", "12x"],
                                  ["mimo", "--cp_len=-3"], ["mimo", "-f", ""]])
def test_bad_command_lines_are_rejected(argv):
    with pytest.raises(rub.RubError):
        rub.config_from_args(argv)


def test_gui_device_record():
    rec = {"ID": "dev0", "Serial": "30C51D4", "Adress": "serial=30C51D4", "Type": "b200", "Product": "B210",
           "TX Gain": 30.0, "RX Gain": 20.5, "Center Freq.": 2.4e9, "Samp. Rate": 1e6,
           "Number of Subcarriers": 512, "Number of Nullcarriers": 40, "Prefix Length": 36.0,
           "Training Sequences": 4.0, "Subdevice Specifications": "A:A A:B"}
    cfg, fe = rub.config_from_json(json.dumps(rec, indent=2))
    assert (cfg.M, cfg.cp_len, cfg.nac) == (512, 36, 4)
    assert fe.num_nullcarriers == 40 and fe.txgain == 30.0 and fe.rxgain == 20.5
    assert fe.cent_freq == 2.4e9 and fe.samp_rate == 1e6
    assert fe.rx_addr == b"serial=30C51D4" and fe.tx_subdev == b"A:A A:B"
    # absent keys leave the defaults alone
    cfg2, fe2 = rub.config_from_json('{"Prefix Length": 18}', cfg=rub.Config(M=256, cp_len=20, num_access_codes=3))
    assert (cfg2.M, cfg2.cp_len, cfg2.nac) == (256, 18, 3)
    with pytest.raises(rub.RubError):
        rub.config_from_json('{"Number of Subcarriers": "many"}')
    with pytest.raises(rub.RubError):
        rub.config_from_json('{"Prefix Length": 1.5}')
