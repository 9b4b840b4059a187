// csrc/host/cli.cc -- `oaxaca-cli` front-end for the mean decomposition, same flags as the reference's clap
// RunArgs (main.rs:44-128) and the same dispatch as run_mean_analysis (main.rs:175-232).  Analysis types other
// than `mean` (Machado-Mata quantile, AKM, matching) are outside the B200 bootstrap path and are refused.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "builder.h"

static std::vector<std::string> split_commas(const std::string& s) {
    std::vector<std::string> out; std::stringstream ss(s); std::string t;
    while (std::getline(ss, t, ',')) if (!t.empty()) out.push_back(t);
    return out;
}

static void usage() {
    std::cerr <<
        "Usage: oaxaca-cli --data <DATA> --outcome <OUTCOME> --group <GROUP> --reference <REFERENCE> [OPTIONS]\n\n"
        "Options:\n"
        "  -d, --data <DATA>                  Path to the input CSV data file\n"
        "      --outcome <OUTCOME>            Outcome variable column\n"
        "      --group <GROUP>                Column that divides the data into two groups\n"
        "      --reference <REFERENCE>        Value of the group column identifying the reference group\n"
        "      --predictors <A,B,..>          Numerical predictor columns\n"
        "      --categorical <A,B,..>         Categorical predictor columns\n"
        "      --analysis-type <TYPE>         mean (default) | quantile | akm | match  [only mean runs on the B200 path]\n"
        "      --ref-coeffs <KIND>            group-a | group-b (default) | pooled | weighted\n"
        "      --bootstrap-reps <N>           Bootstrap replications [default: 50]\n"
        "      --formula <FORMULA>            R-style formula, e.g. \"wage ~ education + C(sector)\"\n"
        "      --weights <COLUMN>             Sample-weight column (WLS)\n"
        "      --normalize <A,B,..>           Categorical variables to Yun-normalise (builder.normalize)\n"
        "      --rif-quantile <TAU>           RIF-regression decomposition at quantile TAU (decompose_quantile)\n"
        "      --seed <N>                     Resampling seed\n"
        "      --output-json <PATH>           Export results as JSON\n"
        "      --output-markdown <PATH>       Export results as Markdown\n";
}

int main(int argc, char** argv) {
    std::map<std::string, std::string> a;
    a["analysis-type"] = "mean"; a["ref-coeffs"] = "group-b"; a["bootstrap-reps"] = "50";   // main.rs:70-83
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        if (k == "-h" || k == "--help") { usage(); return 0; }
        if (k == "-d") k = "--data";
        if (k.rfind("--", 0) != 0 || i + 1 >= argc) { std::cerr << "Error: unexpected argument '" << k << "'\n\n"; usage(); return 2; }
        a[k.substr(2)] = argv[++i];
    }
    try {
        for (const char* req : {"data", "outcome", "group", "reference"})
            if (!a.count(req) && !(std::string(req) == "outcome" && a.count("formula")))
                throw ob::OaxacaError(OB_ERR_INVALID_ARG, std::string("the following required argument was not provided: --") + req);
        if (a["analysis-type"] != "mean")
            throw ob::OaxacaError(OB_ERR_UNSUPPORTED, "analysis type '" + a["analysis-type"] +
                                  "' is outside the B200 bootstrap path (mean decomposition and RIF quantiles only)");
        ob::DataFrame df = ob::DataFrame::read_csv(a["data"]);
        ob::OaxacaBuilder b = a.count("formula")
            ? ob::OaxacaBuilder::from_formula(df, a["formula"], a["group"], a["reference"])
            : ob::OaxacaBuilder(df, a["outcome"], a["group"], a["reference"]);
        if (!a.count("formula")) {
            b.predictors(split_commas(a["predictors"]));
            b.categorical_predictors(split_commas(a["categorical"]));
        }
        const std::string& rc = a["ref-coeffs"];
        b.reference_coefficients(rc == "group-a" ? ob::ReferenceCoefficients::GroupA
                                 : rc == "pooled" ? ob::ReferenceCoefficients::Pooled
                                 : rc == "weighted" ? ob::ReferenceCoefficients::Weighted
                                                    : ob::ReferenceCoefficients::GroupB);
        b.bootstrap_reps((size_t)std::strtoull(a["bootstrap-reps"].c_str(), nullptr, 10));
        if (a.count("weights")) b.weights(a["weights"]);
        if (a.count("normalize")) b.normalize(split_commas(a["normalize"]));
        if (a.count("seed")) b.seed(std::strtoull(a["seed"].c_str(), nullptr, 10));
        if (a.count("selection-outcome")) b.heckman_selection(a["selection-outcome"], split_commas(a["selection-predictors"]));
        const ob::OaxacaResults r = a.count("rif-quantile") ? b.decompose_quantile(std::strtod(a["rif-quantile"].c_str(), nullptr)) : b.run();
        r.summary(std::cout);
        if (a.count("output-json")) { std::ofstream(a["output-json"]) << r.to_json(true, false); }
        if (a.count("output-markdown")) { std::ofstream(a["output-markdown"]) << r.to_markdown(); }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n\n";   // main.rs:379-384: error + help, exit 1
        usage();
        return 1;
    }
}
